#!/usr/bin/env python
"""Where the cycles of chaint_kernel go: wait / work accounting per role thread (MMA issuer, epilogue team, producers, store
warp), from a profiling build of the library (-DFBSNN_CHAIN_PROF, clock64 laps + RED into a global table).

    python tools/chain_prof.py build            # here (CPU box): nvcc -> build/libfbsnn_prof.so
    python tools/chain_prof.py [M] [precision]  # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF_LIB = os.path.join(ROOT, "build", "libfbsnn_prof.so")

SLOTS = {0: "team: wait acc_full", 1: "team: drain (ld, park, arrive)", 2: "team: loop + pick slice / park ld", 3: "team: wait in_full",
         4: "team: math + smem", 5: "team: wait a_free", 6: "team: tcgen05.st + fences", 7: "team: colsum + arrive",
         19: "team: other half-team's chunk (waits)",
         8: "mma: wait acc_empty", 9: "mma: wait a_ready", 10: "mma: wait w_full", 11: "mma: issue + commit",
         12: "store: wait out_ready", 13: "store: TMA store + wait read", 14: "in: wait io_free", 15: "in: issue loads",
         16: "w: wait w_empty", 17: "w: issue loads"}


def build():
    pkg = os.path.join(ROOT, "deep-neural-network-solutions-for-partial-differential-equations_b200")
    sys.path.insert(0, ROOT)
    import importlib
    lib = importlib.import_module("dnnpde_b200")._lib
    os.makedirs(os.path.dirname(PROF_LIB), exist_ok=True)
    cmd = ["nvcc"] + lib.NVCC_FLAGS + ["-DFBSNN_CHAIN_PROF", "-o", PROF_LIB] + lib.SOURCES
    print(" ".join(cmd))
    subprocess.run(cmd, cwd=os.path.join(pkg, "csrc"), check=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        return build()
    os.environ["FBSNN_LIB_PATH"] = PROF_LIB
    import numpy as np
    import torch

    import dnnpde_b200 as pde
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
    lib = pde._lib.load()
    raw = ctypes.CDLL(PROF_LIB)
    sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * 50)[None, :], 1.0, M, 50, 100, [101] + 4 * [256] + [1], "FC", "Sine",
                                     precision=prec, brownian="philox", cuda_graph=False)
    sol.begin_training(1e-3)
    loss = torch.zeros(1, device="cuda")
    for _ in range(2):
        sol.training_step(None, None, loss)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (4 * 160 * 24))()
    raw.fbsnn_debug_chain_prof(None, 1)
    sol.training_step(None, None, loss)
    torch.cuda.synchronize()
    rc = raw.fbsnn_debug_chain_prof(buf, 0)
    assert rc == 0
    a = np.frombuffer(buf, dtype=np.uint64).reshape(4, 160, 24).astype(np.float64)
    for sw, name in enumerate("FATB"):
        t = a[sw]
        n = int((t.sum(axis=1) > 0).sum())
        if n == 0:
            continue
        mean = t[:n].mean(axis=0)
        print(f"== sweep {name}: {n} CTAs; mean cycles per CTA and share of the role's total")
        for ks, role in ((list(range(0, 8)) + [19], "team"), (range(8, 12), "mma"), (range(12, 14), "store"), (range(14, 16), "in"),
                         (range(16, 18), "w")):
            tot = mean[list(ks)].sum()
            for k in ks:
                print(f"   {SLOTS[k]:36s} {mean[k] / 1e6:9.3f} Mcyc  {100 * mean[k] / max(tot, 1):5.1f} %")
            print(f"   {role + ' total':36s} {tot / 1e6:9.3f} Mcyc")


if __name__ == "__main__":
    main()
