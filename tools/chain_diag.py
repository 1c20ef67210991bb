"""Bring-up diagnostic for the layer-chained sweep kernels (gemm_chain.cuh): runs one loss/gradient evaluation with the
per-layer launches (chain = 0) and with the chained sweeps (chain = 2) on identical inputs and reports, array by
array, where the two workspaces differ.  One process per configuration (a faulting kernel kills the context):

    python tools/chain_diag.py --precision tf32x3 --paths 40 [--fwd-only] [--layers 101,256,256,256,256,1] [--act Sine]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def ws_array(sol, lib, sp, M, with_grad, name, layer=0):
    off, w = ctypes.c_int64(), ctypes.c_int()
    rc = lib.fbsnn_debug_ws_offset(ctypes.byref(sp), M, int(with_grad), name.encode(), layer, ctypes.byref(off),
                                   ctypes.byref(w))
    if rc != 0:
        return None
    ws = sol._workspace(lib, sp, M, with_grad)
    rows = M * (sp.N + 1)
    flat = ws.view(torch.float32)
    return flat[off.value:off.value + rows * w.value].view(rows, w.value).clone()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32x3")
    ap.add_argument("--paths", type=int, default=40)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--layers", default="")
    ap.add_argument("--act", default="Sine")
    ap.add_argument("--problem", default="bsb")
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--pair", type=int, default=0, help="chain_pair option (cta_group::2 chain kernel when eligible)")
    ap.add_argument("--ta", type=int, default=1, help="chain_ta option (operand of the next MMA in tensor memory)")
    args = ap.parse_args()
    import dnnpde_b200 as pde
    lib = pde._lib.load()
    lib.fbsnn_set_option(b"chain_ta", args.ta)
    lib.fbsnn_set_option(b"chain_pair", args.pair)
    D, M, N = args.dim, args.paths, args.steps
    layers = [int(x) for x in args.layers.split(",")] if args.layers else [D + 1] + 4 * [256] + [1]
    torch.manual_seed(1)
    np.random.seed(1)
    if args.problem == "hjb":
        sol = pde.HamiltonJacobiBellman(np.zeros((1, D)), 1.0, M, N, D, layers, "FC", args.act, precision=args.precision)
    else:
        sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * (D // 2))[None, :], 1.0, M, N, D, layers, "FC", args.act,
                                         precision=args.precision)
    t, W = sol.fetch_minibatch()
    sp = sol._spec()
    L = len(layers) - 2
    res = {}
    with_grad = not args.fwd_only
    for mode in (0, 2):
        lib.fbsnn_set_option(b"chain", mode)
        if with_grad:
            loss, X, Y, Z, g = sol.loss_grad_flat(t, W, want_Z=True)
            g = g.clone()
        else:
            loss, X, Y, Z = sol._evaluate(t, W, sol.Xi, with_grad=False, want_Z=True)
            g = None
        torch.cuda.synchronize()
        arrs = {"Y": Y.clone(), "Z": Z.clone(), "loss": loss.clone()}
        if g is not None:
            for name, p in sol.model.named_parameters():
                o = sol._fp.offsets[name]
                arrs["grad:" + name] = g[o:o + p.numel()].clone()
        names = ["zf"] + (["V", "ybar"] if with_grad else [])
        for nm in names:
            arrs["ws:" + nm] = ws_array(sol, lib, sp, M, with_grad, nm)
        per_layer = ["g", "a", "delta"] + (["szz", "hd"] if with_grad else [])
        for l in range(1, L + 1):
            for nm in per_layer:
                arrs[f"ws:{nm}[{l}]"] = ws_array(sol, lib, sp, M, with_grad, nm, l)
        res[mode] = arrs
        print(f"mode {mode}: loss {float(loss):.8e}", flush=True)
    if hasattr(lib, "fbsnn_debug_chain_trap") or os.environ.get("FBSNN_LIB_PATH"):
        try:
            raw = ctypes.CDLL(os.environ["FBSNN_LIB_PATH"])
            rec = (ctypes.c_uint * 8)()
            raw.fbsnn_debug_chain_trap(rec, 1)
            print("TRAP RECORD [set, block, role, what, link, chunk, parity, thread]:", list(rec), flush=True)
        except Exception as e:   # noqa: BLE001
            print("no trap record:", e)
    bad = 0
    for k in res[0]:
        a0, a2 = res[0][k], res[2][k]
        if a0 is None or a2 is None:
            continue
        if k in (f"ws:hd[{L}]", f"ws:delta[{L}]") and False:
            continue
        a0, a2 = a0.double().flatten(), a2.double().flatten()
        den = float(a0.abs().max()) + 1e-30
        err = float((a0 - a2).abs().max()) / den
        nanc = int(torch.isnan(a2).sum())
        note = ""
        if k == f"ws:hd[{L}]":
            note = "  (not produced by the chained T sweep: expected to differ)"
        flag = "ok " if err < (2e-2 if args.precision == "tf32" else 2e-5) and nanc == 0 else "BAD"
        if flag == "BAD" and not note:
            bad += 1
            d = (a0 - a2).abs()
            idx = int(d.argmax())
            w = res[0][k].shape[-1] if res[0][k].dim() > 1 else 1
            note += f"  worst at flat {idx} (row {idx // w}, col {idx % w}): {float(a0[idx]):.6e} vs {float(a2[idx]):.6e}"
        print(f"{flag} {k:24s} rel_max_err {err:.3e} nan {nanc}{note}", flush=True)
        if flag == "BAD" and k.startswith("ws:g[") and res[0][k].dim() == 2:
            d2 = (res[0][k].double() - res[2][k].double()).abs() > 1e-4 * den
            rows = torch.nonzero(d2.any(dim=1)).flatten()
            print(f"     {int(rows.numel())} bad rows of {d2.shape[0]}; tiles {sorted(set((rows // 128).tolist()))[:12]}")
            for r in rows[:6].tolist():
                blocks = sorted(set((torch.nonzero(d2[r]).flatten() // 16).tolist()))
                print(f"     row {r} (tile {r // 128}, row-in-tile {r % 128}): bad 16-col blocks {blocks}")
    print("DIAG", "FAIL" if bad else "PASS", args.precision, "fwd-only" if args.fwd_only else "full", f"M={M} layers={layers}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
