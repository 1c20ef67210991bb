#!/usr/bin/env python
"""Per-launch table of the dense-layer GEMMs of one training iteration, timed with CUDA events on the launching
stream (no profiler): sweep tag, ms, algorithmic GB/s and TFLOP/s.  python tools/launch_table.py [M] [precision]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dnnpde_b200 as pde

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
D, N = 100, 50
lib = pde._lib.load()
torch.manual_seed(0)
sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * 50)[None, :], 1.0, M, N, D, [101] + 4 * [256] + [1], "FC", "Sine",
                                 precision=prec, brownian="philox")
sol.begin_training(1e-3)
loss = torch.zeros(1, device="cuda")
for _ in range(2):
    sol.training_step(None, None, loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.fbsnn_dense_timing(1)
e0.record()
sol.training_step(None, None, loss)
e1.record()
torch.cuda.synchronize()
print(f"M={M} precision={prec}: step {e0.elapsed_time(e1):.3f} ms (eager, with timing events)")
out = (ctypes.c_double * 4)()
i, tot = 0, 0.0
while True:
    tag = lib.fbsnn_dense_timing_entry(i, out)
    if tag is None:
        break
    ms, fl, by, tc = out[0], out[1], out[2], out[3]
    tot += ms
    print(f"{i:3d} {tag.decode():10s} {'tc' if tc else 'simt'} {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s  {fl / ms / 1e9:8.1f} TFLOP/s")
    i += 1
print(f"dense total {tot:.3f} ms")
lib.fbsnn_dense_timing(0)
