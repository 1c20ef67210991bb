"""Thin host-side callers of the hot path that the reference's executors use:
the two-phase trainer and the prediction sampler.  Same constructor / method names and return values as upstream
(with_corr_high_dimension_pde.py:619-660 and :704-726); no plotting."""
from __future__ import annotations

import time

import numpy as np
import torch


class TrainingPhases:
    """Initial phase + fine-tuning phase = two `model.train(...)` calls (each builds a fresh Adam, as upstream)."""

    def __init__(self, model):
        self.model = model
        self.min_loss = None
        self.min_loss_state = None

    def _run(self, label, n_iter, lr, optimizer_type):
        print(f"Starting {label} phase...")
        tot = time.time()
        print(self.model.device)
        out = self.model.train(n_iter, lr, optimizer_type)
        print(f"{label.capitalize()} phase completed. Total time:", time.time() - tot, "s")
        if isinstance(out, tuple) and len(out) >= 3:
            self.min_loss, self.min_loss_state = out[1], out[2]
        return out

    def train_initial_phase(self, n_iter, lr, optimizer_type='Adam'):
        return self._run("initial training", n_iter, lr, optimizer_type)

    def fine_tuning_phase(self, n_iter, lr, optimizer_type='Adam'):
        return self._run("fine-tuning", n_iter, lr, optimizer_type)


class PredictionGenerator:
    """`num_samples` fresh minibatches (NumPy stream re-seeded with 42, as upstream) pushed through predict().
    Rows are independent, so all samples go through ONE batched forward on the device; the returned arrays equal
    the upstream per-sample loop's concatenation."""

    def __init__(self, model, Xi, num_samples):
        self.model = model
        self.Xi = Xi
        self.num_samples = num_samples

    def generate_predictions(self):
        np.random.seed(42)
        M0 = self.model.M
        ts, Ws = [], []
        for _ in range(self.num_samples):
            self.model.M = M0
            t_i, W_i = self.model.fetch_minibatch()
            ts.append(t_i), Ws.append(W_i)
        t_all, W_all = torch.cat(ts, 0), torch.cat(Ws, 0)
        X_pred, Y_pred = self.model.predict(self.Xi, t_all, W_all)
        self.model.M = M0                                   # upstream leaves M at the per-sample batch size
        W_test = Ws[0]                                      # upstream returns the first draw's W_test (:629, :660)
        return (t_all.cpu().numpy(), W_test, X_pred.detach().cpu().numpy(), Y_pred.detach().cpu().numpy())
