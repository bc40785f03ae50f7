"""csrc/host_rng.cu against numpy itself: the native replay of the legacy global stream must give the very same numbers
and leave the very same state behind, draw for draw (CPU only; no GPU involved)."""
import numpy as np
import pytest

from dbsgym_b200 import _capi
from dbsgym_b200.np_stream import NumpyGlobalStream


@pytest.fixture(scope="module")
def stream():
    _capi.build()
    return NumpyGlobalStream()


def _same_state():
    a = np.random.get_state()
    return a[1].copy(), a[2], a[3], a[4]


@pytest.mark.parametrize("seed", [0, 10, 12345])
def test_gaussians_and_choices_interleave_exactly_like_numpy(stream, seed):
    np.random.seed(seed)
    ref = [np.random.randn(7), np.random.normal(np.pi, 0.6, 512), np.array([np.random.choice([-1, 1]) for _ in range(40)]),
           np.random.randn(3), np.array([np.random.choice(15) for _ in range(60)]),
           np.array([np.random.choice([4, 5, 6]) for _ in range(60)]), np.random.randn(2 * 512),
           np.array([np.random.choice([3, 4, 5, 6, 7]) for _ in range(50)]), np.random.randn(1), np.random.randn(1000)]
    end = _same_state()
    np.random.seed(seed)
    s = stream.pull()
    got = [s.normal(7), s.normal(512, np.pi, 0.6), np.array([-1, 1])[s.choice(40, 2)], s.normal(3), s.choice(60, 15),
           np.array([4, 5, 6])[s.choice(60, 3)], s.normal(1024), np.array([3, 4, 5, 6, 7])[s.choice(50, 5)], s.normal(1),
           s.normal(1000)]
    s.push()
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    now = _same_state()
    assert np.array_equal(now[0], end[0]) and now[1:] == end[1:]
    # and numpy continues from there as if nothing had happened
    x = np.random.randn(5)
    np.random.seed(seed)
    for r in ref:
        pass
    np.random.set_state(("MT19937", end[0], end[1], end[2], end[3]))
    assert np.array_equal(x, np.random.randn(5))


def test_mixing_native_and_numpy_draws(stream):
    """Half of the draws by numpy, half natively, alternating: same stream as numpy alone (cached Gaussian included)."""
    np.random.seed(77)
    ref = np.concatenate([np.random.randn(3) for _ in range(20)])
    np.random.seed(77)
    got = []
    for k in range(20):
        if k % 2:
            got.append(np.random.randn(3))
        else:
            s = stream.pull(); got.append(s.normal(3)); s.push()
    assert np.array_equal(ref, np.concatenate(got))
