"""Print (not assert) the CUDA-vs-reference deviations for every golden case; run on the GPU box while tuning."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests import golden_util as gu
from tests import parity_util as pu

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
rows = []
for name in gu.solver_cases():
    g, meta = gu.load(name)
    try:
        t0 = time.time()
        sol, oracle = pu.build_cuda_solver(meta, g, precision=prec)
        e = pu.single_eval_errors(sol, oracle, g, meta)
        if not (meta["D"] == 1 and meta["M"] > 1):
            e.update(pu.train_trace_errors(sol, g, meta))
        e["sec"] = round(time.time() - t0, 2)
        print(name, json.dumps({k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in e.items()}), flush=True)
    except Exception as ex:  # noqa
        import traceback
        traceback.print_exc()
        print(name, "FAILED", repr(ex), flush=True)
