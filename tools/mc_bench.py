#!/usr/bin/env python
"""Times the fused MC basket pricer alone (D=100, N=50, correlated): python tools/mc_bench.py [log2 paths]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dnnpde_b200 as pde

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
np.random.seed(0)
model = pde.BlackScholesModel(0.05, 0.2, 100, True)
pr = pde.MonteCarloPricer(model, pde.BasketOption(np.ones(100) / 100, 1.0), 1.0, 50, n, seed=7)
pr.price_async(np.ones(100), n // 16, 0, 7)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sums = pr.price_async(np.ones(100), n, 0, 7)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"variant={os.environ.get('FBSNN_MC_VARIANT', '0')} paths={n} ms={ms:.2f} paths/s={n / ms * 1e3:.4e} price={float(sums[0]) / n:.6f}")
