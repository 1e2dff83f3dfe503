"""CPU ORACLE #2 (test infrastructure, NOT product code): an *independent* restatement of
the reference hot path on ``torch.nn.LSTM`` + autograd, CPU fp32/fp64.

Purpose: (1) cross-check ``oracle/lstm_oracle.py`` (different cell formulation — PyTorch's
(i,f,g,o) fused LSTM — and autograd instead of hand-derived BPTT); (2) serve as the
multi-threaded "reference CPU path" that ``bench.py`` times on the host cores
(`cpu_baseline.kind = "port"`, since TensorFlow 1.x itself cannot be installed here).

Follows reference ``src/models/lstm_baseline.py:38-87`` and SURVEY.md Appendix A.
PARITY UNPINNED (see lstm_oracle.py header).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def _perm_ijfo_to_ifgo(h: int) -> np.ndarray:
    """TF BasicLSTMCell column blocks (i, j, f, o) -> torch row blocks (i, f, g=j, o) (A.2)."""
    idx = np.arange(4 * h).reshape(4, h)
    return np.concatenate([idx[0], idx[2], idx[1], idx[3]])


class TorchRef:
    def __init__(self, params: Dict[str, np.ndarray], config: dict, dtype=torch.float32):
        self.dtype = dtype
        self.scope = config.get("name", "lstm_baseline")
        self.n_layers = int(config.get("n_layers", 1))
        self.h = int(config["hidden_size"])
        self.start_word = int(config["input_size"])
        self.lr = float(config.get("lr", 5e-3))
        self.n_decay = int(config.get("n_decay", 10000))
        self.max_grad_norm = float(config.get("max_grad_norm", 5))
        self.step = 0
        # trainables kept in TF layout/names so Adam state is comparable 1:1
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self._perm = torch.from_numpy(_perm_ijfo_to_ifgo(self.h))

    # -- forward through torch's fused LSTM kernel (weights re-laid-out from TF layout) --
    def _run_lstm(self, inp: torch.Tensor) -> torch.Tensor:
        flat = []
        for layer in range(self.n_layers):
            base = f"{self.scope}/rnn/multi_rnn_cell/cell_{layer}/basic_lstm_cell"
            k, b = self.p[f"{base}/kernel"], self.p[f"{base}/bias"]
            fan = k.shape[0] - self.h
            w_ih = k[:fan, :].index_select(1, self._perm).t().contiguous()
            w_hh = k[fan:, :].index_select(1, self._perm).t().contiguous()
            fb = torch.zeros(4 * self.h, dtype=self.dtype)
            fb[self.h:2 * self.h] = 1.0  # forget_bias=1 added inside the TF cell (A.2)
            b_ih = b.index_select(0, self._perm) + fb
            b_hh = torch.zeros(4 * self.h, dtype=self.dtype)
            flat += [w_ih, w_hh, b_ih, b_hh]
        n = inp.shape[0]
        h0 = torch.zeros(self.n_layers, n, self.h, dtype=self.dtype)
        c0 = torch.zeros_like(h0)
        out, _, _ = torch._VF.lstm(inp, (h0, c0), flat, True, self.n_layers, 0.0, False, False, True)
        return out  # [N,T,H]

    def nll(self, tokens: np.ndarray):
        tok = torch.from_numpy(np.ascontiguousarray(tokens).reshape(-1, tokens.shape[-1]).astype(np.int64))
        x = torch.cat([torch.full((tok.shape[0], 1), self.start_word, dtype=torch.int64), tok[:, :-1]], dim=1)
        emb = self.p[f"{self.scope}/embedding"]
        inp = emb[x]  # IndexedSlices-style gather
        inp.retain_grad() if inp.requires_grad else None
        hs = self._run_lstm(inp)
        logits = hs.reshape(-1, self.h) @ self.p[f"{self.scope}/softmax_w"] + self.p[f"{self.scope}/softmax_b"]
        nll = torch.nn.functional.cross_entropy(logits, tok.reshape(-1), reduction="none")
        return nll.reshape(tok.shape), inp

    def per_token_nll(self, tokens: np.ndarray) -> np.ndarray:
        with torch.no_grad():
            return self.nll(tokens)[0].numpy()

    def eval_loss(self, tokens: np.ndarray) -> float:
        with torch.no_grad():
            nll, _ = self.nll(tokens)
            return float(nll.sum() / (nll.numel() + 1e-12))

    def train_step(self, tokens: np.ndarray) -> float:
        """fwd + autograd bwd + TF clip (A.6) + TF Adam (A.7) + exp-decay (A.8)."""
        for v in self.p.values():
            v.grad = None
        nll, inp = self.nll(tokens)
        loss = nll.sum() / (nll.numel() + 1e-12)
        loss.backward()
        grads = {k: v.grad for k, v in self.p.items()}
        sq = float((inp.grad.double() ** 2).sum())  # un-aggregated IndexedSlices rows
        for k, g in grads.items():
            if not k.endswith("/embedding"):
                sq += float((g.double() ** 2).sum())
        norm = math.sqrt(sq)
        scale = self.max_grad_norm / max(norm, self.max_grad_norm)
        lr_k = float(np.float32(self.lr) * np.power(np.float32(0.5), np.float32(self.step) / np.float32(self.n_decay)))
        t = self.step + 1
        alpha = lr_k * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        with torch.no_grad():
            for k, g in grads.items():
                g = g * scale
                self.m[k].mul_(0.9).add_(g, alpha=0.1)
                self.v[k].mul_(0.999).addcmul_(g, g, value=0.001)
                self.p[k].sub_(alpha * self.m[k] / (self.v[k].sqrt() + 1e-8))
        self.step += 1
        self.last_norm = norm
        return float(loss.detach())

    def params_numpy(self) -> Dict[str, np.ndarray]:
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}
