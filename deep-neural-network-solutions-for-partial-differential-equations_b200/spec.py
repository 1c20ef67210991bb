"""ctypes mirrors of the structs in include/fbsnn_b200.h and the closed enumeration of problem callables."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

MAX_HIDDEN = 8
ABI_VERSION = 103          # fbsnn_version() of the library these struct mirrors were written against
OPT_STATE_BYTES = 2048

NET_FC, NET_NAIS = 0, 1
ACT = {"Sine": 0, "ReLU": 1, "Tanh": 2}
MU_ZERO, MU_LINEAR, MU_HESTON = 0, 1, 2
SIGMA_CONST, SIGMA_PROP, SIGMA_HESTON = 0, 1, 2
PHI_BSB, PHI_RY, PHI_ZSQ = 0, 1, 2
G_SUMSQ, G_CALL_SUM, G_CALL_MEAN, G_LOGQ, G_CALL_FIRST, G_CALL_FIRST_SMOOTH = 0, 1, 2, 3, 4, 5
PRECISION = {"fp32": 0, "tf32": 1, "tf32x3": 2}


class FbsnnSpec(C.Structure):
    _fields_ = [
        ("D", C.c_int32), ("N", C.c_int32), ("n_hidden", C.c_int32),
        ("width", C.c_int32 * MAX_HIDDEN),
        ("net_kind", C.c_int32), ("act_kind", C.c_int32),
        ("mu_kind", C.c_int32), ("sigma_kind", C.c_int32), ("phi_kind", C.c_int32), ("g_kind", C.c_int32),
        ("mu_c", C.c_float), ("sigma_c", C.c_float), ("phi_c", C.c_float), ("strike", C.c_float),
        ("nais_eps", C.c_float),
        ("precision", C.c_int32),
        ("off_W", C.c_int64 * (MAX_HIDDEN + 2)),
        ("off_b", C.c_int64 * (MAX_HIDDEN + 2)),
        ("off_Win", C.c_int64 * (MAX_HIDDEN + 2)),
        ("off_bin", C.c_int64 * (MAX_HIDDEN + 2)),
        ("n_params", C.c_int64),
        ("noise_dim", C.c_int32), ("clamp_u", C.c_int32), ("zt_dims", C.c_int32),
        ("h_kappa", C.c_float), ("h_theta", C.c_float), ("h_xi", C.c_float), ("h_rho", C.c_float), ("h_v0", C.c_float),
    ]


class FbsnnAdam(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("max_grad_norm", C.c_double), ("skip_nonfinite", C.c_double)]


class McSpec(C.Structure):
    _fields_ = [("D", C.c_int32), ("N", C.c_int32), ("rate", C.c_float), ("sigma", C.c_float), ("T", C.c_float),
                ("strike", C.c_float)]


@dataclass(frozen=True)
class ProblemSpec:
    """One row of the problem table in SURVEY.md section 8a: the forms of mu, sigma, phi, g the kernels implement."""
    mu_kind: int
    mu_c: float
    sigma_kind: int
    sigma_c: float
    phi_kind: int
    phi_c: float
    g_kind: int
