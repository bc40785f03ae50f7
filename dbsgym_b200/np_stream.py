"""numpy's legacy global random stream (``np.random.*``), consumed natively for whole batches.

``NumpyGlobalStream`` moves ``np.random.get_state()`` into the C-ABI structure of ``include/dbsgym.h``, lets
``libdbsgym.so`` draw (csrc/host_rng.cu restates MT19937, the polar-method Gaussian and the masked bounded integers of
``RandomState``), and writes the state back, so that draws made here and draws made by numpy interleave exactly as if
numpy had made all of them (reference environment/env.py:291 seeds and uses that global stream for every reset)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


class NumpyGlobalStream:
    def __init__(self):
        self.lib = _capi.load()
        self.st = _capi.DbsGymNpState()

    def pull(self):
        name, key, pos, has_gauss, gauss = np.random.get_state()
        if name != "MT19937":
            raise RuntimeError("np.random global state is not MT19937")
        C.memmove(self.st.key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
        self.st.pos, self.st.has_gauss, self.st.gauss = int(pos), int(has_gauss), float(gauss)
        return self

    def push(self):
        key = np.ctypeslib.as_array(self.st.key).astype(np.uint32)
        np.random.set_state(("MT19937", key, int(self.st.pos), int(self.st.has_gauss), float(self.st.gauss)))

    def normal(self, n, loc=0.0, scale=1.0):
        out = np.empty(int(n))
        rc = self.lib.dbsgym_np_gauss(C.byref(self.st), int(n), float(loc), float(scale), _capi.ptr(out))
        if rc:
            raise _capi.DbsGymError(f"dbsgym_np_gauss failed ({rc})")
        return out

    def choice(self, n, pop_size):
        out = np.empty(int(n), dtype=np.int32)
        rc = self.lib.dbsgym_np_choice(C.byref(self.st), int(n), int(pop_size), _capi.ptr(out))
        if rc:
            raise _capi.DbsGymError(f"dbsgym_np_choice failed ({rc})")
        return out
