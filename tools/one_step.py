#!/usr/bin/env python
"""Two eager training iterations (one warm-up, one to profile) of the bench workload, for ncu captures:
python tools/one_step.py [M] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dnnpde_b200 as pde

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * 50)[None, :], 1.0, M, 50, 100, [101] + 4 * [256] + [1], "FC", "Sine",
                                 precision=prec, brownian="philox", cuda_graph=False)
sol.begin_training(1e-3)
loss = torch.zeros(1, device="cuda")
for _ in range(2):
    sol.training_step(None, None, loss)
torch.cuda.synchronize()
print("loss", float(loss))
