"""FBSNN solver with the reference's Python surface, driven by the sm_100a kernels behind include/fbsnn_b200.h.

Mirrors `FBSNN(ABC)` of the reference (DeepBSDE.py:140-323, with_corr_high_dimension_pde.py:132-543,
1d_BSPDE_case.py:126-508): same constructor arities, attributes, `net_u`, `loss_function`, `fetch_minibatch`,
`train`, `predict`, `save_model`, `load_model`.  Every compute method runs the CUDA path; there is no CPU or
eager-PyTorch fallback -- without a CUDA device or the built library they raise.

Differences from the reference that a user can see (all stated in DESIGN.md):
  * `mu_tf/sigma_tf/phi_tf/g_tf` stay callable in Python, but the kernels evaluate the closed enumeration in
    `problems.py`; a subclass outside it raises NotImplementedError instead of silently running slowly.
  * the N-schedule that silently replaces N is opt-in: `n_schedule="reference"`.
  * for D == 1 and M > 1 the per-path product Z . (sigma dW) is used, not the reference's cross-path broadcast
    (SURVEY section 9 Q3).
  * `net_u` outputs are not differentiable w.r.t. X (the analytic adjoint is evaluated in-kernel).
"""
from __future__ import annotations

import ctypes
import math
import time
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import parallel
from . import spec as S
from .networks import FlatParams, build_model, make_activation

_TORCH_OPTIMS = ("SGD", "RMSprop", "AdamW", "Adadelta", "Adagrad", "Adamax", "ASGD")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _LossWithGrad(torch.autograd.Function):
    """Makes the kernel-computed loss a differentiable scalar: backward hands out the parameter gradients the
    fused step already produced (loss_function(...)[0].backward() then fills p.grad like the reference)."""

    @staticmethod
    def forward(ctx, solver, loss, grads_flat, *params):
        ctx.solver = solver
        ctx.grads = grads_flat
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        fp = ctx.solver._fp
        outs = []
        for name, p in fp.model.named_parameters():
            o = fp.offsets[name]
            outs.append(ctx.grads[o:o + p.numel()].view(p.shape) * gout)
        return (None, None, None, *outs)


class FBSNN(ABC):
    # class-level switches set by the problem subclasses / module aliases
    problem_spec: Optional[S.ProblemSpec] = None
    _strike_per_dim = False          # 1d_/nd_BSPDE_case.py:160 use strike = 1.0 * D
    _y0_as_float = False             # DeepBSDE.loss_function returns Y[0,0,0].item()
    _log_every = 500                 # with_corr...:436 (DeepBSDE.py:284 uses 100)
    _train_returns = "quad"          # "graph" | "triple" | "quad"
    _clip_norm: Optional[float] = 1.0  # with_corr...:424; DeepBSDE has no clipping
    _schedule_kind = "mm"            # "mm" (1d/nd/hjb) | "recursive" (with_corr) -- only with n_schedule="reference"
    _skip_nonfinite = False          # heston_dnnpde.py:408-410: NaN-loss iterations are skipped
    _check_layers = True             # layers[0] == D + 1 (the Heston class swaps its input layer afterwards)

    def __init__(self, Xi, T, M, N, D, *args, **kw):
        # two reference arities: (layers, mode, activation) and (Mm, layers, mode, activation[, correlation_type])
        names_short = ["layers", "mode", "activation"]
        names_long = ["Mm", "layers", "mode", "activation", "correlation_type"]
        extras = {k: kw.pop(k) for k in ("precision", "n_schedule", "brownian", "seed", "device", "data_parallel",
                                          "cuda_graph", "collective") if k in kw}
        if len(args) == 3 or (len(args) < 3 and "Mm" not in kw and len(args) + len(kw) == 3):
            vals = dict(zip(names_short, args))
            self._arity = "short"
        else:
            vals = dict(zip(names_long, args))
            self._arity = "long"
        vals.update(kw)
        unknown = set(vals) - set(names_long)
        if unknown:
            raise TypeError(f"unexpected arguments {sorted(unknown)}")
        layers, mode, activation = vals["layers"], vals["mode"], vals["activation"]
        Mm = vals.get("Mm")
        correlation_type = vals.get("correlation_type", "no_correlation")

        dev = extras.get("device")
        if dev is not None:
            self.device = torch.device(dev)
        elif torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device("cpu")   # host-side logic only; every compute method raises
        Xi = np.asarray(Xi)
        self.Xi = torch.from_numpy(Xi).float().to(self.device)
        self.Xi.requires_grad = True

        self.T, self.M, self.N, self.D, self.Mm = T, M, N, D, Mm
        self.strike = 1.0 * D if self._strike_per_dim else 1.0
        self.mode = mode
        self.activation = activation
        self.activation_function = make_activation(activation)
        if self._check_layers and (list(layers)[0] != D + 1 or list(layers)[-1] != 1):
            raise ValueError(f"layers must start with D+1={D + 1} and end with 1, got {list(layers)}")
        self.layers = list(layers)
        self.model = build_model(self.layers, mode, self.activation_function).to(self.device)
        self.model.apply(self.weights_init)
        self._fp = FlatParams(self.model, "FC" if mode == "FC" else "NAIS", self.device)

        self.training_loss = []
        self.iteration = []
        self.optimizer = None
        self.correlation_type = correlation_type
        self.correlation_matrix = self.generate_correlation_matrix(D)

        # arithmetic variant of the dense layers: "tf32x3" (default; tensor cores, fp32-grade), "fp32" (SIMT FMA,
        # tightest parity), "tf32" (single-pass tensor cores, looser stated tolerance)
        self.precision = extras.get("precision", "tf32x3")
        if self.precision not in S.PRECISION:
            raise ValueError(f"precision {self.precision!r} not in {sorted(S.PRECISION)}")
        self.n_schedule = extras.get("n_schedule")
        if self.n_schedule not in (None, "reference"):
            raise ValueError("n_schedule must be None (fixed N) or 'reference'")
        self.brownian = extras.get("brownian", "numpy")
        if self.brownian not in ("numpy", "philox"):
            raise ValueError("brownian must be 'numpy' (reference stream, host) or 'philox' (in-kernel)")
        self.seed = int(extras.get("seed", 0))
        self.data_parallel = bool(extras.get("data_parallel", False))
        self._ws = {}
        self._chol_dev = None
        self._opt_state = None
        self._n_train_calls = 0
        self._graphs = {}
        self.use_cuda_graph = bool(extras.get("cuda_graph", True))
        self._peer = None
        self.collective = extras.get("collective", "peer")   # multi-GPU: "peer" (fused NVLink kernel) | "nccl"

    @property
    def _sdim(self) -> int:
        """State dimension the kernels see (= D; the Heston class carries (S, v) on one Brownian driver)."""
        return self.D

    def _state_xi(self, Xi: torch.Tensor) -> torch.Tensor:
        """(rows, state dim) initial condition from the caller's Xi."""
        return Xi.reshape(-1, self.D)

    def _fill_problem(self, sp: S.FbsnnSpec) -> None:
        ps = self.problem_spec
        sp.mu_kind, sp.mu_c = ps.mu_kind, ps.mu_c
        sp.sigma_kind, sp.sigma_c = ps.sigma_kind, ps.sigma_c
        sp.phi_kind, sp.phi_c = ps.phi_kind, ps.phi_c
        sp.g_kind, sp.strike = ps.g_kind, self.strike

    # ------------------------------------------------------------------------------------------------
    # host-side pieces kept verbatim in meaning
    # ------------------------------------------------------------------------------------------------
    def weights_init(self, m):
        if type(m) == nn.Linear:
            torch.nn.init.xavier_uniform_(m.weight)

    def generate_correlation_matrix(self, D):
        """with_corr_high_dimension_pde.py:187-195 (identity for the default 'no_correlation')."""
        if self.correlation_type == "no_correlation":
            return np.eye(D)
        if self.correlation_type == "random_correlation":
            return self.generate_random_correlation_matrix(D)
        if self.correlation_type == "restricted_random_correlation":
            return self.generate_random_correlation_matrix(D, restrict_positive=True)
        raise ValueError("Invalid correlation type")

    def generate_random_correlation_matrix(self, D, restrict_positive=False):
        """Reference generator (with_corr...:197-212), drawn from the NumPy global RNG; note that its diagonal is
        not 1 -- reproduced as is, the kernels only ever see its Cholesky factor."""
        a = np.random.randn(D, D)
        if restrict_positive:
            a = np.abs(a)
        c = np.dot(a, a.T)
        np.fill_diagonal(c, 1)
        d = np.sqrt(np.diag(c))
        c = c / np.outer(d, d)
        return self._make_positive_definite(c)

    def _make_positive_definite(self, matrix):
        eps = 1e-6
        while not np.all(np.linalg.eigvals(matrix) > 0):
            matrix += eps * np.eye(matrix.shape[0])
            eps *= 2
        return matrix

    def _cholesky(self) -> Optional[np.ndarray]:
        if self.correlation_type == "no_correlation":
            return None
        return np.linalg.cholesky(self.correlation_matrix)

    def fetch_minibatch(self):
        """Host Brownian sampler on the NumPy global RNG, exactly the reference's stream and layout
        (DeepBSDE.py:247-262; correlated: with_corr...:316-353): t (M,N+1,1), W (M,N+1,D) cumulative, fp32."""
        M, N, D, T = self.M, self.N, self.D, self.T
        Dt = np.zeros((M, N + 1, 1))
        DW = np.zeros((M, N + 1, D))
        dt = T / N
        Dt[:, 1:, :] = dt
        inc = np.sqrt(dt) * np.random.normal(size=(M, N, D))
        L = self._cholesky()
        if L is not None:
            inc = np.einsum('ij,mnj->mni', L, inc)
        DW[:, 1:, :] = inc
        t = torch.from_numpy(np.cumsum(Dt, axis=1)).float().to(self.device)
        W = torch.from_numpy(np.cumsum(DW, axis=1)).float().to(self.device)
        return t, W

    def fetch_minibatch_device(self, seed=None, iteration=0, path_offset=0, n_paths=None, N=None):
        """Device-side twin of fetch_minibatch(): same layout and distribution (incl. the Cholesky-correlated form),
        drawn by Philox4x32-10 keyed (seed, iteration, GLOBAL path id) -- so any contiguous shard
        [path_offset, path_offset + n_paths) reproduces exactly the rows of the full minibatch.  Not NumPy's
        stream."""
        lib = self._require_cuda()
        M = self.M if n_paths is None else int(n_paths)
        N = self.N if N is None else int(N)
        sp = self._spec(N)
        dev = self.device
        with torch.cuda.device(dev):
            ws = self._workspace(lib, sp, M, True)
            t = torch.empty(M, N + 1, 1, device=dev)
            W = torch.empty(M, N + 1, self.D, device=dev)
            chol = self._chol_device()
            rc = lib.fbsnn_fetch_minibatch(ctypes.byref(sp), float(self.T), M, int(path_offset),
                                           int(self.seed if seed is None else seed), int(iteration), _ptr(chol),
                                           _ptr(ws), ws.numel(), _ptr(t), _ptr(W), self._stream())
            _lib.check(rc, "fbsnn_fetch_minibatch")
        return t, W

    # ------------------------------------------------------------------------------------------------
    # kernel plumbing
    # ------------------------------------------------------------------------------------------------
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("FBSNN compute needs a CUDA device (sm_100a); there is no CPU fallback")
        if self.problem_spec is None:
            raise NotImplementedError(
                f"{type(self).__name__}: mu/sigma/phi/g are not in the closed enumeration the fused kernels "
                "implement (set `problem_spec`); refusing to fall back to eager PyTorch")
        if not self._fp.is_intact():
            raise RuntimeError("model parameters no longer alias the flat buffer (was the model moved or cast?)")
        return _lib.load()

    def _spec(self, N: Optional[int] = None) -> S.FbsnnSpec:
        sp = S.FbsnnSpec()
        sp.D, sp.N = self._sdim, self.N if N is None else N
        sp.net_kind = S.NET_FC if self.mode == "FC" else S.NET_NAIS
        sp.act_kind = S.ACT[self.activation]
        self._fill_problem(sp)
        sp.nais_eps = 0.01
        sp.precision = S.PRECISION[self.precision]
        self._fp.fill_spec(sp)
        return sp

    def _workspace(self, lib, sp, n_paths: int, with_grad: bool) -> torch.Tensor:
        need = ctypes.c_size_t(0)
        _lib.check(lib.fbsnn_workspace_bytes(ctypes.byref(sp), n_paths, int(with_grad), ctypes.byref(need)),
                   "fbsnn_workspace_bytes")
        key = "grad" if with_grad else "fwd"
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need.value:
            self._ws[key] = None
            self._graphs.clear()            # captured graphs hold the old workspace address
            ws = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        off = (-ws.data_ptr()) % 256
        return ws[off:]

    def _chol_device(self) -> Optional[torch.Tensor]:
        L = self._cholesky()
        if L is None:
            return None
        if self._chol_dev is None:
            self._chol_dev = torch.from_numpy(np.ascontiguousarray(L)).float().to(self.device)
        return self._chol_dev

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _as_f32(x, device, shape=None):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.asarray(x))
        x = x.detach().to(device=device, dtype=torch.float32).contiguous()
        return x if shape is None else x.reshape(shape)

    # ------------------------------------------------------------------------------------------------
    # reference surface
    # ------------------------------------------------------------------------------------------------
    def net_u(self, t, X):
        """u(t, X) (M,1) and Du = du/dX (M,D) (DeepBSDE.py:189-194), analytic adjoint in-kernel."""
        lib = self._require_cuda()
        t = self._as_f32(t, self.device)
        X = self._as_f32(X, self.device)
        if t.dim() == 1:
            t = t.unsqueeze(-1)
        if X.dim() == 1:
            X = X.unsqueeze(-1)
        rows = X.shape[0]
        sp = self._spec()
        with torch.cuda.device(self.device):
            ws = self._workspace(lib, sp, max(1, math.ceil(rows / (sp.N + 1))), False)
            u = torch.empty(rows, 1, device=self.device)
            du = torch.empty(rows, self._sdim, device=self.device)
            _lib.check(lib.fbsnn_net_u(ctypes.byref(sp), _ptr(self._fp.flat), _ptr(t), _ptr(X), rows, _ptr(ws),
                                       ws.numel(), _ptr(u), _ptr(du), self._stream()), "fbsnn_net_u")
        return u, du

    def Dg_tf(self, X):
        """Gradient of the terminal condition by autograd on the Python callable (DeepBSDE.py:196-200); the
        fused loss uses the closed form instead."""
        X = X.detach().requires_grad_(True)
        g = self.g_tf(X)
        return torch.autograd.grad(g, X, grad_outputs=torch.ones_like(g), allow_unused=True)[0]

    def _evaluate(self, t, W, Xi, with_grad: bool, want_Z: bool = False):
        lib = self._require_cuda()
        t = self._as_f32(t, self.device)
        W = self._as_f32(W, self.device)
        M = W.shape[0]
        N = W.shape[1] - 1
        if N != self.N:
            raise ValueError(f"W has {N} steps but self.N = {self.N}")
        if W.shape[2] != self.D or t.shape[:2] != W.shape[:2]:
            raise ValueError(f"bad shapes t{tuple(t.shape)} W{tuple(W.shape)} for D={self.D}")
        Xi = self._state_xi(self._as_f32(Xi, self.device)).contiguous()
        if Xi.shape[0] not in (1, M):
            raise ValueError(f"Xi has {Xi.shape[0]} rows, expected 1 or {M}")
        sp = self._spec()
        dev = self.device
        with torch.cuda.device(dev):
            ws = self._workspace(lib, sp, M, with_grad)
            X = torch.empty(M, N + 1, self._sdim, device=dev)
            Y = torch.empty(M, N + 1, 1, device=dev)
            Z = torch.empty(M, N + 1, self._sdim, device=dev) if want_Z else None
            loss = torch.empty((), device=dev)
            if with_grad:
                rc = lib.fbsnn_loss_grad(ctypes.byref(sp), _ptr(self._fp.flat), _ptr(self._fp.grad), _ptr(t), _ptr(W),
                                         _ptr(Xi), Xi.shape[0], M, float(self.T), 0, 0, 0, None, _ptr(ws), ws.numel(),
                                         _ptr(X), _ptr(Y), _ptr(Z), _ptr(loss), self._stream())
                _lib.check(rc, "fbsnn_loss_grad")
            else:
                rc = lib.fbsnn_forward(ctypes.byref(sp), _ptr(self._fp.flat), _ptr(t), _ptr(W), _ptr(Xi), Xi.shape[0],
                                       M, _ptr(ws), ws.numel(), _ptr(X), _ptr(Y), _ptr(Z), _ptr(loss), self._stream())
                _lib.check(rc, "fbsnn_forward")
        return loss, X, Y, Z

    def loss_function(self, t, W, Xi):
        """(loss, X, Y, Y0) as DeepBSDE.py:202-245.  `loss.backward()` fills p.grad of every model parameter with
        the analytically computed gradient (written out in-kernel, not by autograd)."""
        loss, X, Y, _ = self._evaluate(t, W, Xi, with_grad=True)
        grads = self._fp.grad.clone()
        params = [p for _, p in self.model.named_parameters()]
        for p in params:
            p.requires_grad_(True)
        loss_t = _LossWithGrad.apply(self, loss, grads, *params)
        y0 = Y[0, 0, 0]
        return loss_t, X, Y, (y0.item() if self._y0_as_float else y0)

    def loss_grad_flat(self, t, W, Xi=None, want_Z=False):
        """Kernel-level entry used by the parity tests: (loss, X, Y, Z, flat gradient buffer)."""
        loss, X, Y, Z = self._evaluate(t, W, self.Xi if Xi is None else Xi, with_grad=True, want_Z=want_Z)
        return loss, X, Y, Z, self._fp.grad

    def predict(self, Xi_star, t_star, W_star):
        """(X_star, Y_star) (DeepBSDE.py:297-302; with_corr...:455-486).  Accepts NumPy or tensors, broadcasts
        singleton batch dimensions and, like the reference, sets self.M to the batch size."""
        Xi_star = self._as_f32(Xi_star, self.device)
        Xi_star = Xi_star.reshape(-1, Xi_star.shape[-1] if Xi_star.dim() > 1 else self.D)
        t_star = self._as_f32(t_star, self.device)
        W_star = self._as_f32(W_star, self.device)
        batch = max(Xi_star.shape[0], t_star.shape[0], W_star.shape[0])
        self.M = batch
        if t_star.shape[0] == 1:
            t_star = t_star.repeat(batch, 1, 1)
        if W_star.shape[0] == 1:
            W_star = W_star.repeat(batch, 1, 1)
        _, X, Y, _ = self._evaluate(t_star, W_star, Xi_star, with_grad=False)
        return X, Y

    # ------------------------------------------------------------------------------------------------
    # training
    # ------------------------------------------------------------------------------------------------
    def _scheduled_N(self, it: int) -> int:
        """The reference's N-schedule, only applied with n_schedule='reference'."""
        if self._schedule_kind == "recursive":          # with_corr...:406-409
            if 4000 <= it < 20000:
                return int(np.ceil((self.N ** (1 / 5)) ** (int(it / 4000) + 1)))
            if it < 4000:
                return int(np.ceil(self.N ** (1 / 5)))
            return self.N
        if self.Mm is None:                             # hjb_implement.py:591 passes None -> TypeError upstream (Q4)
            return self.N
        if 4000 <= it < 20000:                          # 1d_BSPDE_case.py:372-375
            return int(np.ceil(self.Mm ** (int(it / 4000) + 1)))
        if it < 4000:
            return int(np.ceil(self.Mm))
        return self.N

    def _shard(self):
        if self.data_parallel and parallel.is_distributed():
            return parallel.shard_range(self.M, parallel.rank(), parallel.world_size())
        return 0, self.M

    def train(self, N_Iter, learning_rate, optimizer_type="Adam"):
        """Training loop of the reference (DeepBSDE.py:265-295, with_corr...:355-453): a fresh Adam per call,
        one Brownian minibatch per iteration, summed-residual loss, analytic backward, optional clip, Adam.
        The whole iteration runs on the device; the loss is read back only where the reference logs it."""
        if optimizer_type != "Adam":
            if optimizer_type in _TORCH_OPTIMS:
                return self._train_generic(N_Iter, learning_rate, optimizer_type)
            if optimizer_type == "LBFGS":
                raise NotImplementedError("LBFGS needs closure re-evaluation; not on the fused path")
            raise ValueError(f"Optimizer type '{optimizer_type}' is not recognized.")
        self._require_cuda()
        dev = self.device
        previous_it = self.iteration[-1] if self.iteration else 0
        self.begin_training(learning_rate)
        track_min = self._train_returns not in ("graph", "heston")
        loss_buf = torch.zeros(N_Iter + 1, device=dev)
        y0_buf = torch.zeros(N_Iter + 1, device=dev)
        time_logs = []
        cumulative, start = 0.0, time.time()
        last_logged = 0
        with torch.cuda.device(dev):
            track = self._begin_min_tracking() if track_min else None
            for k, it in enumerate(range(previous_it, previous_it + N_Iter)):
                if self.n_schedule == "reference":
                    n_new = self._scheduled_N(it)
                    if n_new != self.N:
                        self.N = n_new
                        if track is not None:        # the trajectories change shape: keep the best so far, re-arm
                            self._flush_min_tracking(track)
                if self.brownian == "numpy":
                    t_b, W_b = self.fetch_minibatch()
                else:
                    t_b = W_b = None
                # the loss, Y0 and the min-loss bookkeeping (with_corr...:428-433) all stay on the device: nothing is
                # read back between the reference's logging points
                X, Y = self._step(t_b, W_b, loss_buf[k:k + 1], track is not None and track["keep_X"], k, track=track)
                y0_buf[k] = Y[0, 0, 0]
                if it % self._log_every == 0:
                    vals = loss_buf[last_logged:k + 1].cpu().numpy()
                    last_logged = k + 1
                    elapsed = time.time() - start
                    cumulative += elapsed
                    time_logs.append(cumulative)
                    if self._skip_nonfinite and not np.isfinite(vals[-1]):
                        # heston_dnnpde.py:408-410 `continue`s on a NaN loss: that iteration is neither averaged nor logged
                        last_logged = k + 1 - int(np.isfinite(vals).sum())   # keep the finite ones for the next window
                        start = time.time()
                        continue
                    if self._arity == "short":
                        print('It: %d, Loss: %.3e, Y0: %.3f, Time: %.2f, Learning Rate: %.3e' %
                              (it, float(vals[-1]), float(y0_buf[k]), elapsed, learning_rate))
                    start = time.time()
                    if self._skip_nonfinite:
                        vals = vals[np.isfinite(vals)]
                    self.training_loss.append(vals.mean())
                    self.iteration.append(it)
                    if self._train_returns == "heston":
                        self.Y0_values.append(float(y0_buf[k]))
            self.last_losses = loss_buf[:N_Iter].cpu().numpy()     # device -> host read of the step results
            self.last_Y0 = y0_buf[:N_Iter].cpu().numpy()
            min_loss, min_loss_state = self._finish_min_tracking(track) if track is not None else (float("inf"), None)
        # p.grad = the gradient the last Adam step used (summed over ranks on the multi-GPU peer path)
        self._fp.attach_grads(self._peer["sum"][:self._fp.n] if self._peer is not None else None)
        if self._train_returns == "heston":               # heston_dnnpde.py:448
            return np.column_stack((self.iteration, self.training_loss, self.Y0_values))
        graph = np.stack((self.iteration, self.training_loss))
        if self._train_returns == "graph":
            return graph
        if self._train_returns == "triple":
            return graph, min_loss, min_loss_state
        return graph, min_loss, min_loss_state, time_logs

    # ------------------------------------------------------------------------------------------------
    # min_loss / min_loss_state (with_corr...:431-433) without a per-iteration host sync
    # ------------------------------------------------------------------------------------------------
    def _begin_min_tracking(self):
        """Device-side record of the best iteration: fbsnn_track_min keeps (best loss, its iteration index) and copies
        the trajectories of an improving iteration on the device.  With in-kernel Brownian increments X does not depend
        on the parameters, so only Y (M x (N+1) floats) is copied and X is re-materialised once, at the end, from the
        Philox (seed, iteration) of the best step; with host-supplied increments X is copied too."""
        state = torch.zeros(8, device=self.device)
        self._reset_track_state(state)
        return {"state": state, "keep_X": self.brownian == "numpy", "Xb": None, "Yb": None, "N": None,
                "best": (float("inf"), None), "shard": self._shard()}

    @staticmethod
    def _reset_track_state(state):
        state[0] = float("inf")
        state[1:4].view(torch.int32).copy_(torch.tensor([0, -1, 0], dtype=torch.int32))

    def _track_buffers(self, track, X, Y):
        if track["N"] != self.N or track["Yb"] is None:
            track["N"] = self.N
            track["Yb"] = torch.zeros_like(Y)
            track["Xb"] = torch.zeros_like(X) if (track["keep_X"] and X is not None) else None
        return track["Xb"], track["Yb"]

    def _enqueue_track(self, lib, track, loss, X, Y):
        Xb, Yb = self._track_buffers(track, X, Y)
        n_x = X.numel() if Xb is not None else 0
        if n_x % 4 or Y.numel() % 4:              # odd sizes: host-free fallback with torch ops on the device
            better = loss.reshape(()) < track["state"][0]
            track["state"][0] = torch.where(better, loss.reshape(()), track["state"][0])
            Yb.copy_(torch.where(better, Y, Yb))
            if Xb is not None:
                Xb.copy_(torch.where(better, X, Xb))
            ints = track["state"][1:4].view(torch.int32)
            ints[1] = torch.where(better, ints[2], ints[1])
            ints[2] += 1
            it64 = track["state"][4:6].view(torch.int64)
            it64[0] = torch.where(better, self._opt_state[24:32].view(torch.int64)[0] - 1, it64[0])
            return
        _lib.check(lib.fbsnn_track_min(_ptr(loss), _ptr(track["state"]), _ptr(self._opt_state), _ptr(X) if n_x else None,
                                       _ptr(Xb) if n_x else None, n_x, _ptr(Y), _ptr(Yb), Y.numel(), self._stream()),
                   "fbsnn_track_min")

    def _flush_min_tracking(self, track):
        """Fold the device record into the running best (called when N changes, and at the end of train())."""
        if track["Yb"] is None:
            return
        host = track["state"].cpu()
        best = float(host[0])
        k_best = int(host[1:4].view(torch.int32)[1])
        rng_best = int(host[4:6].view(torch.int64)[0])
        if k_best >= 0 and best < track["best"][0]:
            Yb = track["Yb"].clone()
            if track["keep_X"]:
                Xb = track["Xb"].clone()
            else:   # X of the best step from its Philox stream: (seed, device iteration counter at that step)
                Xb = self._paths_philox(rng_best, track["N"], *track["shard"])
            track["best"] = (best, (Xb.detach(), Yb.detach()))
        self._reset_track_state(track["state"])
        track["Yb"] = None

    def _finish_min_tracking(self, track):
        self._flush_min_tracking(track)
        return track["best"]

    def _paths_philox(self, iteration, N, lo, hi):
        """X (m_loc, N+1, D) of the training step that drew its increments with Philox `iteration`: X does not depend
        on the parameters, so it is re-materialised exactly by one more evaluation of that step's kernels (gradients
        into a scratch buffer) instead of having been copied on every improving iteration."""
        lib = self._require_cuda()
        fp, dev = self._fp, self.device
        sp = self._spec(N)
        m_loc = hi - lo
        ws = self._workspace(lib, sp, m_loc, True)
        X = torch.empty(m_loc, N + 1, self._sdim, device=dev)
        scratch = torch.empty(fp.n, device=dev)
        loss = torch.empty((), device=dev)
        Xi = self._state_xi(self.Xi.detach())
        xi_loc = Xi.contiguous() if Xi.shape[0] == 1 else Xi[lo:hi].contiguous()
        chol = self._chol_device()
        rc = lib.fbsnn_loss_grad(ctypes.byref(sp), _ptr(fp.flat), _ptr(scratch), None, None, _ptr(xi_loc), xi_loc.shape[0],
                                 m_loc, float(self.T), lo, self.seed & 0xFFFFFFFFFFFFFFFF, int(iteration), _ptr(chol),
                                 _ptr(ws), ws.numel(), _ptr(X), None, None, _ptr(loss), self._stream())
        _lib.check(rc, "fbsnn_loss_grad")
        return X

    def begin_training(self, learning_rate):
        """Fresh Adam state, as the reference builds a new optimiser in every train() call."""
        fp = self._fp
        fp.exp_avg.zero_(), fp.exp_avg_sq.zero_()
        if self._opt_state is None:
            self._opt_state = torch.zeros(S.OPT_STATE_BYTES, dtype=torch.uint8, device=self.device)
        self._opt_state[:24].zero_()       # Adam step counter; the Philox iteration counter at byte 24 lives on
        self.optimizer = {"type": "Adam", "lr": learning_rate, "betas": (0.9, 0.999), "eps": 1e-8}
        self._hp = S.FbsnnAdam(learning_rate, 0.9, 0.999, 1e-8, self._clip_norm if self._clip_norm else 0.0,
                               1.0 if self._skip_nonfinite else 0.0)
        self._n_train_calls += 1

    _MAX_GRAPHS = 6

    def _step(self, t_b, W_b, loss_out, want_X, k, alias_inputs=False, track=None):
        """training_step through a captured CUDA graph when possible: the ~20-40 kernels of an iteration are replayed
        with one launch, which is what bounds small-M steps.  The Adam step and Philox iteration counters live on the
        device, so a replay is a genuinely new iteration.  Multi-GPU steps are captured too when the gradient exchange
        is the library's own peer-memory kernel (its cross-GPU barrier is keyed by the same device counter); the NCCL
        exchange stays eager.  `track` = device-side min-loss bookkeeping enqueued right after the step (in the graph)."""
        lib = self._require_cuda()
        dp = self.data_parallel and parallel.is_distributed()
        if dp and self.collective == "peer":
            self._peer_buffers()                  # may fall back to "nccl" (agreed across ranks)
        if not self.use_cuda_graph or (dp and self.collective != "peer"):
            X, Y = self.training_step(t_b, W_b, loss_out, want_X=want_X)
            if track is not None:
                self._enqueue_track(lib, track, loss_out, X, Y)
            return X, Y
        host_batch = W_b is not None
        # everything a captured kernel argument depends on: shapes, hyper-parameters baked in by value (lr, clip, T, seed)
        # and the addresses of tensors the caller may replace (Xi, the aliased minibatch, the tracking buffers)
        key = (self.M, self.N, self.precision, bool(want_X), host_batch, self._hp.lr, self._hp.max_grad_norm,
               self.seed, float(self.T), self.Xi.data_ptr(), id(track["state"]) if track is not None else None,
               (t_b.data_ptr(), W_b.data_ptr()) if (alias_inputs and host_batch) else None)
        gs = self._graphs.get(key)
        if gs is None:
            dev = self.device
            while len(self._graphs) >= self._MAX_GRAPHS:      # bounded: an N-schedule or many lr values would otherwise
                self._graphs.pop(next(iter(self._graphs)))    # pin one set of X / Y outputs per key for ever
            gs = {"loss": torch.zeros(1, device=dev)}
            if host_batch and alias_inputs:     # caller keeps (t_b, W_b) alive and in place: read them directly
                gs["t"], gs["W"] = t_b, W_b
            elif host_batch:
                gs["t"], gs["W"] = torch.empty_like(t_b), torch.empty_like(W_b)
                gs["t"].copy_(t_b), gs["W"].copy_(W_b)
            # capture needs a warm-up run on a side stream; keep it side-effect free by restoring the state (Adam step,
            # Philox iteration counter).  The one thing that is NOT rewound is the epoch of the peer all-reduce's cross-GPU
            # barrier (bytes 40..48 of the optimiser state): every rank advances it alike and it must never run backwards
            fp = self._fp
            saved = [x.clone() for x in (fp.flat, fp.exp_avg, fp.exp_avg_sq, self._opt_state)]
            tsaved = track["state"].clone() if track is not None else None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                Xw, Yw = self.training_step(gs.get("t"), gs.get("W"), gs["loss"], want_X=want_X)
                if track is not None:
                    self._track_buffers(track, Xw, Yw)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gs["X"], gs["Y"] = self.training_step(gs.get("t"), gs.get("W"), gs["loss"], want_X=want_X)
                if track is not None:
                    self._enqueue_track(lib, track, gs["loss"], gs["X"], gs["Y"])
                    gs["track_N"] = track["N"]
            for dst, src in zip((fp.flat, fp.exp_avg, fp.exp_avg_sq), saved):
                dst.copy_(src)
            if dp:
                epoch_now = self._opt_state[40:48].clone()
                self._opt_state.copy_(saved[3])
                self._opt_state[40:48].copy_(epoch_now)
            else:
                self._opt_state.copy_(saved[3])
            if track is not None:
                track["state"].copy_(tsaved)
            gs["graph"] = graph
            self._graphs[key] = gs
        if track is not None and (track["Yb"] is None or track["N"] != self.N):
            # tracking buffers were re-armed (N changed and came back): the captured copies point at the old ones
            self._graphs.pop(key, None)
            return self._step(t_b, W_b, loss_out, want_X, k, alias_inputs, track)
        if host_batch and not alias_inputs:
            gs["t"].copy_(t_b), gs["W"].copy_(W_b)
        gs["graph"].replay()
        loss_out.copy_(gs["loss"])
        return gs["X"], gs["Y"]

    def training_step(self, t_b, W_b, loss_out, want_X=False, iteration=None):
        """One training iteration, enqueued on the current stream without synchronising.  (t_b, W_b) are the
        GLOBAL minibatch in the reference layout, or None to draw the Brownian increments in-kernel (Philox).
        With data_parallel and torch.distributed initialised each rank evaluates its contiguous slice of the
        paths and the [gradient | loss] buffer is all-reduced before the (replicated) Adam update."""
        lib = self._require_cuda()
        fp, dev = self._fp, self.device
        dp = self.data_parallel and parallel.is_distributed()
        sp = self._spec()
        lo, hi = self._shard()
        m_loc = hi - lo
        ws = self._workspace(lib, sp, m_loc, True)
        X = torch.empty(m_loc, self.N + 1, self._sdim, device=dev) if want_X else None
        Y = torch.empty(m_loc, self.N + 1, 1, device=dev)
        if W_b is not None:
            if W_b.shape[0] == self.M and m_loc != self.M:
                t_b, W_b = t_b[lo:hi], W_b[lo:hi]
            t_b, W_b = t_b.contiguous(), W_b.contiguous()
            if W_b.shape != (m_loc, self.N + 1, self.D):
                raise ValueError(f"minibatch shape {tuple(W_b.shape)} != {(m_loc, self.N + 1, self.D)}")
            chol = None
        else:
            chol = self._chol_device()
        Xi = self._state_xi(self.Xi.detach())
        xi_loc = Xi.contiguous() if Xi.shape[0] == 1 else Xi[lo:hi].contiguous()
        seed = self.seed & 0xFFFFFFFFFFFFFFFF
        if not dp:
            rc = lib.fbsnn_train_step(ctypes.byref(sp), ctypes.byref(self._hp), _ptr(fp.flat), _ptr(fp.grad),
                                      _ptr(fp.exp_avg), _ptr(fp.exp_avg_sq), _ptr(self._opt_state), _ptr(t_b),
                                      _ptr(W_b), _ptr(xi_loc), xi_loc.shape[0], m_loc, float(self.T), lo, seed, 0,
                                      _ptr(chol), _ptr(ws), ws.numel(), _ptr(X), _ptr(Y), _ptr(loss_out),
                                      self._stream())
            _lib.check(rc, "fbsnn_train_step")
        else:
            peer = self._peer_buffers() if self.collective == "peer" else None
            loss_slot = loss_out
            if peer:
                # peers have finished reading the previous iteration's gradients from this rank's buffer
                loss_slot = peer["buf"][fp.n:fp.n + 1]
                _lib.check(lib.fbsnn_peer_wait(_ptr(peer["buf"]), fp.n, parallel.world_size(), _ptr(self._opt_state),
                                               self._stream()), "fbsnn_peer_wait")
            # the Philox iteration is the persistent device counter of the optimiser state -- the same one the single-GPU
            # fbsnn_train_step uses -- so the sharded step draws exactly the 1-GPU stream and consecutive train() calls
            # (TrainingPhases) never replay noise
            rc = lib.fbsnn_loss_grad_step(ctypes.byref(sp), _ptr(fp.flat), _ptr(fp.grad), _ptr(t_b), _ptr(W_b),
                                          _ptr(xi_loc), xi_loc.shape[0], m_loc, float(self.T), lo, seed, _ptr(chol),
                                          _ptr(self._opt_state), _ptr(ws), ws.numel(), _ptr(X), _ptr(Y),
                                          _ptr(loss_slot), self._stream())
            _lib.check(rc, "fbsnn_loss_grad_step")
            if peer:
                # one kernel: cross-GPU barrier + sum of the peers' buffers over NVLink + clip norm; then Adam
                rc = lib.fbsnn_peer_allreduce_adam(ctypes.byref(self._hp), _ptr(fp.flat),
                                                   ctypes.c_void_p(peer["ptrs"]), parallel.world_size(), parallel.rank(),
                                                   _ptr(peer["sum"]), _ptr(fp.exp_avg), _ptr(fp.exp_avg_sq), fp.n,
                                                   _ptr(self._opt_state), self._stream())
                _lib.check(rc, "fbsnn_peer_allreduce_adam")
                loss_out.copy_(peer["sum"][fp.n:fp.n + 1])
            else:
                parallel.allreduce_grads_and_loss(fp.grad, loss_out)
                rc = lib.fbsnn_adam_step(ctypes.byref(self._hp), _ptr(fp.flat), _ptr(fp.grad), _ptr(fp.exp_avg),
                                         _ptr(fp.exp_avg_sq), fp.n, _ptr(self._opt_state), self._stream())
                _lib.check(rc, "fbsnn_adam_step")
        return X, Y

    def _peer_buffers(self):
        """Symmetric (peer-mapped) [gradient | loss | flags] buffer for the fused NVLink all-reduce
        (fbsnn_peer_allreduce_adam).  torch's symmetric-memory allocator only provides the peer mapping; the
        barrier and the reduction are this library's kernel.  Falls back to NCCL if the devices cannot map each
        other's memory."""
        if self._peer is not None or self.collective != "peer":
            return self._peer
        import torch.distributed as dist
        peer, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm
            lib = _lib.load()
            fp = self._fp
            flag_off, total = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(lib.fbsnn_peer_buffer_floats(fp.n, ctypes.byref(flag_off), ctypes.byref(total)),
                       "fbsnn_peer_buffer_floats")
            buf = symm.empty(total.value, dtype=torch.float32, device=self.device)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            peer = {"buf": buf, "hdl": hdl, "ptrs": int(hdl.buffer_ptrs_dev),
                    "sum": torch.zeros(fp.n + 4, device=self.device)}
        except Exception as e:                 # noqa: BLE001 -- e.g. no P2P between the devices
            err = e
        # the choice between the peer kernel and NCCL must be the same on every rank (a rank that fell back alone would
        # wait in an NCCL all-reduce for peers that spin in the peer kernel): agree on the minimum
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # also: every rank has zeroed its flags before anyone can signal
        if int(ok) == 1:
            self._peer = peer
            self._fp.grad = peer["buf"][:self._fp.n]   # the gradient kernels now write straight into the shared buffer
            self._fp.attach_grads()
        else:
            import warnings
            warnings.warn(f"symmetric memory unavailable on at least one rank ({err!r}); using the NCCL all-reduce")
            self.collective = "nccl"
            self._peer = None
        return self._peer

    def _train_generic(self, N_Iter, learning_rate, optimizer_type):
        """Non-Adam optimisers of the reference's menu (with_corr...:370-389): gradients from the fused kernels,
        update by torch.optim on the flat parameter views."""
        import torch.optim as optim
        self._require_cuda()
        previous_it = self.iteration[-1] if self.iteration else 0
        self.optimizer = getattr(optim, optimizer_type)(self.model.parameters(), lr=learning_rate)
        loss_temp = []
        min_loss, min_loss_state, time_logs = float("inf"), None, []
        for it in range(previous_it, previous_it + N_Iter):
            if self.n_schedule == "reference":
                self.N = self._scheduled_N(it)
            t_b, W_b = self.fetch_minibatch()
            loss, X, Y, _ = self._evaluate(t_b, W_b, self.Xi, with_grad=True)
            self._fp.attach_grads()
            if self._clip_norm:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self._clip_norm)
            self.optimizer.step()
            lv = float(loss)
            loss_temp.append(lv)
            if lv < min_loss:
                min_loss, min_loss_state = lv, (X.clone(), Y.clone())
            if it % self._log_every == 0:
                self.training_loss.append(float(np.mean(loss_temp)))
                loss_temp = []
                self.iteration.append(it)
        graph = np.stack((self.iteration, self.training_loss))
        if self._train_returns == "graph":
            return graph
        if self._train_returns == "triple":
            return graph, min_loss, min_loss_state
        return graph, min_loss, min_loss_state, time_logs

    # ------------------------------------------------------------------------------------------------
    # checkpoints (with_corr...:488-499): same dict keys; optimiser state is not saved, as upstream
    # ------------------------------------------------------------------------------------------------
    def save_model(self, file_name):
        torch.save({'model_state_dict': self.model.state_dict(), 'training_loss': self.training_loss,
                    'iteration': self.iteration}, file_name)

    def load_model(self, file_name):
        checkpoint = torch.load(file_name, map_location=self.device, weights_only=False)
        self.model.load_state_dict(checkpoint['model_state_dict'])
        self.training_loss = checkpoint['training_loss']
        self.iteration = checkpoint['iteration']

    # ------------------------------------------------------------------------------------------------
    # problem callables (Python versions stay available for drivers and plots)
    # ------------------------------------------------------------------------------------------------
    @abstractmethod
    def phi_tf(self, t, X, Y, Z):
        pass

    @abstractmethod
    def g_tf(self, X):
        pass

    @abstractmethod
    def mu_tf(self, t, X, Y, Z):
        return torch.zeros([self.M, self.D]).to(self.device)

    @abstractmethod
    def sigma_tf(self, t, X, Y):
        return torch.diag_embed(torch.ones([self.M, self.D])).to(self.device)
