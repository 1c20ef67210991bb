"""Network containers with the reference's module / state_dict naming, and the flat parameter layout.

The modules hold the parameters (so `model.state_dict()`, `load_state_dict`, `parameters()` and
`model.apply(weights_init)` behave as in the reference); the FBSNN hot path never calls their `forward` --
it reads the parameters through one flat fp32 buffer that every `nn.Parameter` is a view of.

Reference: Functions/Sine.py:6-12, Functions/naisnet.py:6-95 (inline copy with_corr_high_dimension_pde.py:29-129),
FC construction DeepBSDE.py:166-172, xavier init DeepBSDE.py:185-187.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import spec as S

ALIGN = 32  # floats: every tensor starts on a 128-byte boundary of the flat buffer


class Sine(nn.Module):
    """sin activation as a module (Functions/Sine.py:6-12)."""

    def forward(self, x):
        return torch.sin(x)


def make_activation(name: str) -> nn.Module:
    if name == "Sine":
        return Sine()
    if name == "ReLU":
        return nn.ReLU()
    if name == "Tanh":
        return nn.Tanh()
    raise ValueError(f"activation {name!r} is not one of 'Sine', 'ReLU', 'Tanh'")


class Naisnet(nn.Module):
    """NAIS-Net parameter container, 1..3 stable blocks for len(layers) in {4, 5, 6}.

    Attribute names equal the reference's (layer1, layer2, layer2_input, ...), so checkpoints interchange.
    `forward` is a plain-torch convenience for callers that evaluate `model.model(x)` directly (the reference's
    analytics do); unlike the reference it builds the identity on the weight's device.
    """

    def __init__(self, layers, stable: bool, activation: nn.Module):
        super().__init__()
        if len(layers) not in (4, 5, 6):
            raise ValueError("Naisnet supports len(layers) in {4, 5, 6}")
        self.layers = list(layers)
        self.layer1 = nn.Linear(layers[0], layers[1])
        self.layer2 = nn.Linear(layers[1], layers[2])
        self.layer2_input = nn.Linear(layers[0], layers[2])
        self.layer3 = nn.Linear(layers[2], layers[3])
        if len(layers) >= 5:
            self.layer3_input = nn.Linear(layers[0], layers[3])
            self.layer4 = nn.Linear(layers[3], layers[4])
        if len(layers) == 6:
            self.layer4_input = nn.Linear(layers[0], layers[4])
            self.layer5 = nn.Linear(layers[4], layers[5])
        self.activation = activation
        self.epsilon = 0.01
        self.stable = stable

    @property
    def n_blocks(self) -> int:
        return len(self.layers) - 3

    def project(self, layer, out):
        w = layer.weight
        delta = 1 - 2 * self.epsilon
        rtr = w.t() @ w
        norm = torch.norm(rtr)
        if norm > delta:
            rtr = delta ** 0.5 * rtr / (norm ** 0.5)
        a = rtr + torch.eye(rtr.shape[0], device=w.device, dtype=w.dtype) * self.epsilon
        return F.linear(out, -a, layer.bias)

    def forward(self, x):
        u = x
        out = self.activation(self.layer1(x))
        for k in range(2, self.n_blocks + 2):
            lin = getattr(self, f"layer{k}")
            shortcut = out
            if self.stable:
                out = self.project(lin, out) + getattr(self, f"layer{k}_input")(u)
            else:
                out = lin(out)
            out = self.activation(out) + shortcut
        return getattr(self, f"layer{self.n_blocks + 2}")(out)


def build_model(layers, mode: str, activation_module: nn.Module) -> nn.Module:
    if mode == "FC":
        mods: List[nn.Module] = []
        for i in range(len(layers) - 2):
            mods.append(nn.Linear(in_features=layers[i], out_features=layers[i + 1]))
            mods.append(activation_module)
        mods.append(nn.Linear(in_features=layers[-2], out_features=layers[-1]))
        return nn.Sequential(*mods)
    if mode in ("Naisnet", "NAIS-Net"):
        return Naisnet(layers, stable=True, activation=activation_module)
    raise NotImplementedError(
        f"mode {mode!r}: only 'FC' and 'Naisnet' ('NAIS-Net') have fused sm_100a kernels (no CPU fallback); "
        "Resnet/Verlet/SDEnet are never selected by the reference's drivers")


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def layer_table(model: nn.Module, mode: str) -> Tuple[list, int]:
    """[(l, W, b, Win, bin)] for hidden layers l = 1..L plus the output layer L+1 (Win/bin None when absent)."""
    rows = []
    if mode == "FC":
        lins = [m for m in model if isinstance(m, nn.Linear)]
        for l, lin in enumerate(lins, start=1):
            rows.append((l, lin.weight, lin.bias, None, None))
        return rows, len(lins) - 1
    nb = model.n_blocks
    rows.append((1, model.layer1.weight, model.layer1.bias, None, None))
    for k in range(2, nb + 2):
        lin, inp = getattr(model, f"layer{k}"), getattr(model, f"layer{k}_input")
        rows.append((k, lin.weight, lin.bias, inp.weight, inp.bias))
    out = getattr(model, f"layer{nb + 2}")
    rows.append((nb + 2, out.weight, out.bias, None, None))
    return rows, nb + 1


class FlatParams:
    """Re-homes every parameter of `model` into one flat fp32 buffer (plus equally laid out gradient and Adam
    buffers) and records the offsets the kernels need."""

    def __init__(self, model: nn.Module, mode: str, device: torch.device):
        self.model = model
        self.mode = mode
        self.device = device
        named = list(model.named_parameters())
        offs, off = {}, 0
        for name, p in named:
            offs[name] = off
            off = _round_up(off + p.numel(), ALIGN)
        self.n = max(off, ALIGN)
        self.offsets = offs
        self.flat = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._by_param = {}
        with torch.no_grad():
            for name, p in named:
                o = offs[name]
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                self._by_param[id(p)] = o
        self.attach_grads()

    def offset_of(self, p) -> int:
        return self._by_param[id(p)]

    def attach_grads(self, src=None):
        """Point every p.grad at its slice of the flat gradient buffer (or of `src`, same layout: the reduced
        gradients of the multi-GPU peer all-reduce)."""
        buf = self.grad if src is None else src
        for name, p in self.model.named_parameters():
            o = self.offsets[name]
            p.grad = buf[o:o + p.numel()].view(p.shape)

    def grad_views(self):
        return [self.grad[self.offsets[n]:self.offsets[n] + p.numel()].view(p.shape)
                for n, p in self.model.named_parameters()]

    def is_intact(self) -> bool:
        """True while every parameter still aliases the flat buffer (e.g. not after model.to(other_device))."""
        base = self.flat.data_ptr()
        for name, p in self.model.named_parameters():
            if p.data_ptr() != base + 4 * self.offsets[name] or p.dtype != torch.float32:
                return False
        return True

    def fill_spec(self, sp: S.FbsnnSpec):
        rows, L = layer_table(self.model, self.mode)
        for arr in (sp.off_W, sp.off_b, sp.off_Win, sp.off_bin):
            for i in range(S.MAX_HIDDEN + 2):
                arr[i] = -1
        for l, W, b, Win, bin_ in rows:
            sp.off_W[l] = self.offset_of(W)
            sp.off_b[l] = self.offset_of(b)
            if Win is not None:
                sp.off_Win[l] = self.offset_of(Win)
                sp.off_bin[l] = self.offset_of(bin_)
        sp.n_params = self.n
        sp.n_hidden = L
        for i in range(S.MAX_HIDDEN):
            sp.width[i] = 0
        for l, W, *_ in rows[:L]:
            sp.width[l - 1] = W.shape[0]
        return L
