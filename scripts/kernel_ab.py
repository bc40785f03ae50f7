#!/usr/bin/env python
"""A/B timing of the float32 step kernels on one GPU (device-resident, CUDA events inside the library):
exact multi-worker contraction vs the spectral kernels (spectral64 = 64-thread workers, warp34 / warp32 = one warp per
environment with 34 / 32 modes), same environments, same actions."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--modes", default="exact,spectral64,warp34,warp32")
    args = ap.parse_args()
    import torch
    from dbsgym_b200.batched import BatchedKuramoto
    B = args.envs
    dicts = build_params(B)
    rng = np.random.default_rng(0)
    acts = torch.from_numpy(rng.uniform(-1, 1, (args.steps + 5, B)).astype(np.float32)).cuda()
    ref = None
    for mode in args.modes.split(","):
        np.random.seed(0)
        t0 = time.time()
        kw = {"exact": dict(coupling_eval="exact"),
              "spectral64": dict(coupling_eval="spectral", engine_options={"no_warp_kernel": True}),
              "warp34": dict(coupling_eval="spectral", spectral_tol=1e-10),
              "warp32": dict(coupling_eval="spectral", spectral_tol=1e-9)}[mode]
        core = BatchedKuramoto([dict(d) for d in dicts], **kw)
        eng = core.engine
        eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
        eng.set_timing(True)
        ms = []
        for k in range(args.steps + 5):
            eng.step_device(acts[k].data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            if k >= 5:
                ms.append(eng.last_step_ms()[0])
        y = eng.state()
        c = eng.counters()
        line = f"{mode:10s} modes {(eng.spectral or {}).get('modes')} variant {eng.step_variant()} step kernel {np.mean(ms):.4f} ms (min {np.min(ms):.4f}) -> " \
               f"{B / np.mean(ms) * 1e3 / 1e6:.2f} M env-steps/s; status {c['status']}; setup {time.time() - t0:.1f}s"
        if ref is not None:
            line += f"; max |phase - {args.modes.split(',')[0]}| after {args.steps + 5} free-running steps {np.max(np.abs(y - ref)):.2e}"
        else:
            ref = y
        print(line, flush=True)
        core.close()


if __name__ == "__main__":
    main()
