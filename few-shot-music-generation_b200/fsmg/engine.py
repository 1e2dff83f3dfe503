"""Host-side engine: owns the PyTorch tensors (parameters, Adam slots, flat gradient buffer,
workspace) and drives libfsmg.so through its C-ABI.  PyTorch is plumbing here (device memory,
streams, torch.distributed); all arithmetic of the hot path runs in the CUDA library.

Replaces the TensorFlow session of the reference (src/models/tf_model.py:80-97 and the
sess.run calls of src/models/lstm_baseline.py:104,125,150).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import FsmgError, fsmg_config, fsmg_param_info


def glorot_uniform_init(shapes: Dict[str, tuple], seed: int) -> Dict[str, np.ndarray]:
    """TF-1.x default initialiser for get_variable (glorot_uniform): every
    trainable U(-l, l), l = sqrt(6/(fan_in+fan_out)) — including softmax_b, with fan_in =
    fan_out = len for vectors — except the LSTM bias, which BasicLSTMCell zero-initialises.
    TF's RNG stream cannot be reproduced, so only the distribution matches
    (reference tf_model.py:81 seeds with config['seed'])."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        if name.endswith("/bias"):
            out[name] = np.zeros(shape, np.float32)
            continue
        fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else shape
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        out[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
    return out


class DeviceLoss:
    """Mean loss of a step, still on the device: `sum_nll` is a 1-element view of the gradient buffer's scalar slot, valid until the
    next step; float() reads it back (synchronising) and divides on the host — no extra kernel in the step."""
    __slots__ = ("sum_nll", "tokens")

    def __init__(self, sum_nll: torch.Tensor, tokens: float):
        self.sum_nll, self.tokens = sum_nll, tokens

    def __float__(self) -> float:
        return float(self.sum_nll) / (self.tokens + 1e-12)

    item = __float__


class Engine:
    """One engine per process/GPU.  `config` uses the reference's keys (lstm_baseline.py:21-29)."""

    def __init__(self, config: dict, max_seqs: int, device: Optional[torch.device] = None, flags: int = 0,
                 process_group=None, world: Optional[int] = None):
        if not torch.cuda.is_available():
            raise FsmgError("fsmg requires a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self.config = dict(config)
        self.V = int(config["input_size"])
        self.E = int(config["embedding_size"])
        self.H = int(config["hidden_size"])
        self.L = int(config.get("n_layers", 1))
        self.T = int(config["max_len"])
        self.max_seqs = int(max_seqs)
        self.scope = str(config.get("name", "lstm_baseline"))
        flags |= int(os.environ.get("FSMG_FLAGS", "0"))
        cfg = fsmg_config(
            vocab=self.V, embed=self.E, hidden=self.H, layers=self.L, max_len=self.T, max_seqs=self.max_seqs,
            n_decay=int(config.get("n_decay", 10000)), flags=flags, lr=float(config.get("lr", 5e-3)),
            max_grad_norm=float(config.get("max_grad_norm", 5)), beta1=0.9, beta2=0.999, eps=1e-8, reserved=0.0)
        self.flags = flags
        h = C.c_void_p()
        _lib.check(self.lib.fsmg_create(C.byref(cfg), self.scope.encode(), C.byref(h)))
        self.h = h
        self.n_params = int(self.lib.fsmg_param_count(h))
        self.n_grads = int(self.lib.fsmg_grad_count(h))
        ws_bytes = int(self.lib.fsmg_workspace_bytes(h))
        dev = self.device
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.n_grads, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.workspace = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
        _lib.check(self.lib.fsmg_bind(h, self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                      self.adam_v.data_ptr(), ws_ptr, ws_bytes))
        self.infos: List[dict] = []
        for i in range(self.lib.fsmg_num_params(h)):
            pi = fsmg_param_info()
            _lib.check(self.lib.fsmg_param_info_at(h, i, C.byref(pi)))
            shape = (pi.rows, pi.cols) if pi.cols > 1 or not pi.name.decode().endswith(("/bias", "/softmax_b")) else (pi.rows,)
            self.infos.append(dict(name=pi.name.decode(), offset=int(pi.offset), shape=shape))
        self.global_step = 0
        self.pg = process_group
        # world = 1 inside an initialised process group gives a purely local engine (single-GPU cross-checks of a DP run)
        self.world = int(world) if world is not None else (torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1)
        self._nll = torch.empty(self.max_seqs * self.T, dtype=torch.float32, device=dev)
        self._sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self._tok = torch.empty(self.max_seqs * self.T, dtype=torch.int32, device=dev)
        self._pinned_tok = torch.empty(self.max_seqs * self.T, dtype=torch.int32).pin_memory()
        self._pinned_scal = torch.empty(8, dtype=torch.float32).pin_memory()
        self._h2d_done: Optional[torch.cuda.Event] = None
        self._ids: Optional[torch.Tensor] = None
        self._pinned_np = None
        self._checked_corpora = set()
        self._overlap = False
        # single GPU: the loss of a step is final ~40 % into it.  It is read back from a side stream behind an event recorded inside
        # the step graph, so train_host*() returns while the backward pass and the update still run and the caller's next episode is
        # sampled / staged concurrently; later calls are stream-ordered behind the step.  FSMG_EARLY_LOSS=0: wait for the whole step.
        self._early_loss = False
        if self.world == 1 and os.environ.get("FSMG_EARLY_LOSS", "1") != "0":
            self._loss_stream = torch.cuda.Stream(device=self.device)
            self._ev_loss = torch.cuda.Event()
            self._ev_loss.record(torch.cuda.current_stream(self.device))      # materialises the cudaEvent_t
            _lib.check(self.lib.fsmg_set_loss_event(self.h, self._ev_loss.cuda_event))
            self._early_loss = True
        # Off by default: measured on 2 and 8 B200 (profiles/r2_optimization_log.md) the overlapped schedule is within noise of the
        # single all-reduce — the collective is ~0.3 ms of an 11 ms step and what N > 1 loses is mostly cross-GPU skew.
        if self.world > 1 and os.environ.get("FSMG_AR_OVERLAP", "0") == "1":
            self._setup_overlapped_allreduce()

    # ---- data-parallel gradient all-reduce overlapped with the backward pass ----------------------------------------
    AR_CTAS = 4      # CTAs of the side communicator's kernels = SMs the persistent recurrent kernels leave free

    def _setup_overlapped_allreduce(self) -> None:
        """The softmax_w | softmax_b slice of the flat gradient buffer (41 % of it at configs[1]) is final when the projection
        backward ends, ~3 ms before the step does: it is all-reduced from a stage event (recorded inside the library's CUDA
        graph) on a side stream, through a communicator limited to AR_CTAS CTAs that fit the SMs the persistent recurrent
        kernels leave free.  Embedding | LSTM kernels | biases (one contiguous range) and the 8 scalars follow after the call at
        full width.  Measured on 8 B200 (profiles/r2_optimization_log.md): also reducing the embedding slice on the 4-CTA
        side communicator from a second event was slower than reducing it at full width afterwards."""
        import torch.distributed as dist
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = self.AR_CTAS
            opts.config.min_ctas = 1
            self._pg_side = dist.new_group(backend="nccl", pg_options=opts)
        except Exception:           # no NCCL config support: keep the single all-reduce
            return
        self._ranges = []
        for which in range(4):
            b, e = C.c_int64(), C.c_int64()
            _lib.check(self.lib.fsmg_param_range(self.h, which, C.byref(b), C.byref(e)))
            self._ranges.append((int(b.value), int(e.value)))
        self._side = torch.cuda.Stream(device=self.device)
        self._ev_soft, self._ev_emb = torch.cuda.Event(), torch.cuda.Event()
        for ev in (self._ev_soft, self._ev_emb):
            ev.record(torch.cuda.current_stream(self.device))     # materialises the cudaEvent_t
        _lib.check(self.lib.fsmg_set_stage_events(self.h, self._ev_soft.cuda_event, None, self.AR_CTAS))
        # warm both communicators up (lazy NCCL init would otherwise land inside the first step)
        probe = torch.zeros(8, device=self.device)
        dist.all_reduce(probe, group=self._pg_side)
        dist.all_reduce(probe, group=self.pg)
        torch.cuda.synchronize(self.device)
        self._overlap = True

    def _all_reduce_grads(self) -> None:
        import torch.distributed as dist
        if not self._overlap:
            dist.all_reduce(self.grads, group=self.pg)
            return
        main = torch.cuda.current_stream(self.device)
        g, (emb, rnn, soft, extra) = self.grads, self._ranges
        assert emb[1] == rnn[0] and rnn[1] == soft[0] and soft[1] == extra[0]
        self._side.wait_event(self._ev_soft)
        with torch.cuda.stream(self._side):
            dist.all_reduce(g[soft[0]:soft[1]], group=self._pg_side)
        dist.all_reduce(g[emb[0]:rnn[1]], group=self.pg)          # after the whole backward pass, on the caller's stream
        dist.all_reduce(g[extra[0]:extra[1]], group=self.pg)
        main.wait_stream(self._side)

    # ---- parameters ---------------------------------------------------------------------------
    def param_shapes(self) -> Dict[str, tuple]:
        return {i["name"]: tuple(i["shape"]) for i in self.infos}

    def _view(self, flat: torch.Tensor, info: dict) -> torch.Tensor:
        n = int(np.prod(info["shape"]))
        return flat[info["offset"]: info["offset"] + n].view(*info["shape"])

    def param_views(self, which: str = "params") -> Dict[str, torch.Tensor]:
        flat = {"params": self.params, "grads": self.grads, "adam_m": self.adam_m, "adam_v": self.adam_v}[which]
        return {i["name"]: self._view(flat, i) for i in self.infos}

    def load_params(self, params: Dict[str, np.ndarray], strict: bool = True) -> List[str]:
        """Inject weights by TF variable name (name AND shape must match, like
        optimistic_restore, reference tf_model.py:28-75).  Returns the names loaded."""
        loaded = []
        views = self.param_views()
        for name, view in views.items():
            if name not in params:
                if strict:
                    raise KeyError(name)
                continue
            arr = np.asarray(params[name], dtype=np.float32)
            if tuple(arr.shape) != tuple(view.shape):
                if strict:
                    raise ValueError(f"{name}: shape {arr.shape} != {tuple(view.shape)}")
                continue
            view.copy_(torch.from_numpy(arr))
            loaded.append(name)
        self.refresh_weights()
        return loaded

    def init_params(self, seed: int) -> None:
        self.load_params(glorot_uniform_init(self.param_shapes(), seed))

    def export(self, which: str = "params") -> Dict[str, np.ndarray]:
        return {k: v.detach().cpu().numpy().copy() for k, v in self.param_views(which).items()}

    def refresh_weights(self) -> None:
        _lib.check(self.lib.fsmg_refresh_weights(self.h, self._stream()))

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- device-resident entry points ---------------------------------------------------------------
    def forward_nll(self, tokens: torch.Tensor):
        """tokens int32 [N,T] on the device -> (per-token nll [N,T] fp32, sum nll [1])."""
        n = int(tokens.shape[0])
        assert tokens.dtype == torch.int32 and tokens.is_cuda and tokens.is_contiguous() and tokens.shape[1] == self.T
        nll = self._nll[: n * self.T]
        _lib.check(self.lib.fsmg_forward_nll(self.h, tokens.data_ptr(), n, nll.data_ptr(), self._sum.data_ptr(), self._stream()))
        return nll.view(n, self.T), self._sum

    def forward_backward(self, tokens: torch.Tensor, global_tokens: Optional[int] = None) -> None:
        n = int(tokens.shape[0])
        assert tokens.dtype == torch.int32 and tokens.is_cuda and tokens.is_contiguous() and tokens.shape[1] == self.T
        gt = global_tokens if global_tokens is not None else n * self.T * self.world
        loss_scale = float(1.0 / (float(gt) + 1e-12))
        _lib.check(self.lib.fsmg_forward_backward(self.h, tokens.data_ptr(), n, loss_scale, 0, self._stream()))

    def train_step_device(self, tokens: torch.Tensor, global_tokens: Optional[int] = None) -> DeviceLoss:
        """One optimizer step on device-resident tokens.  Data parallel: every rank passes its
        shard; ONE all-reduce (sum) of the flat gradient buffer — which also carries sum(nll)
        and the per-occurrence embedding-gradient square norm — then clip+Adam on every rank.
        Returns the mean loss of the global batch as a DeviceLoss (read back with float())."""
        n = int(tokens.shape[0])
        gt = global_tokens if global_tokens is not None else n * self.T * self.world
        self.forward_backward(tokens, gt)
        if self.world > 1:
            self._all_reduce_grads()
        _lib.check(self.lib.fsmg_apply_update(self.h, self.global_step, 0, self._stream()))
        self.global_step += 1
        self._last_gt = float(gt)
        self._last_gt_assumed = global_tokens is None
        return DeviceLoss(self.grads[self.n_params: self.n_params + 1], float(gt))

    def _read_loss(self) -> float:
        """D2H of the step's scalars (sum of NLL, token count) after train_step_device: the mean loss of the global batch.  The
        all-reduced token count (grads[n_params + 2], written by the library) must equal the count the loss scale assumed —
        unequal shards across ranks would otherwise train with a silently wrong gradient scale."""
        if self._early_loss:
            self._loss_stream.wait_event(self._ev_loss)
            with torch.cuda.stream(self._loss_stream):
                self._pinned_scal[:4].copy_(self.grads[self.n_params: self.n_params + 4], non_blocking=True)
            self._loss_stream.synchronize()
        else:
            self._pinned_scal[:4].copy_(self.grads[self.n_params: self.n_params + 4], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        total, count = float(self._pinned_scal[0]), float(self._pinned_scal[2])
        if self._last_gt_assumed and abs(count - self._last_gt) > 1e-6 * self._last_gt:
            raise FsmgError(f"ranks processed {count:.0f} tokens in this step but the loss was scaled for {self._last_gt:.0f}: "
                            "every rank must hold the same number of sequences, or pass global_tokens explicitly")
        return total / (self._last_gt + 1e-12)

    def sample_greedy_device(self, n_songs: int, n_tokens: int) -> torch.Tensor:
        out = torch.empty((n_songs, n_tokens), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.fsmg_sample_greedy(self.h, n_songs, n_tokens, out.data_ptr(), self._stream()))
        return out

    # ---- host-buffer entry points (numpy in, python scalars out: what LSTMBaseline uses) -------------
    def _check_ids(self, tok: np.ndarray) -> None:
        """TensorFlow's embedding_lookup raises for ids outside the table (reference lstm_baseline.py:41); so do we —
        a vocabulary / input_size mismatch must not train on clamped garbage."""
        if tok.size and (int(tok.min()) < 0 or int(tok.max()) > self.V):
            raise FsmgError(f"token ids span [{int(tok.min())}, {int(tok.max())}], outside [0, {self.V}] (input_size = {self.V})")

    def _stage(self, tokens: np.ndarray) -> torch.Tensor:
        tok = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1, self.T)
        n = tok.shape[0]
        if n > self.max_seqs:
            raise FsmgError(f"{n} sequences > engine capacity {self.max_seqs}")
        self._check_ids(tok)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()      # the previous asynchronous H2D copy has finished reading the pinned buffer
        self._pinned_tok[: n * self.T].copy_(torch.from_numpy(tok.reshape(-1)))
        dev = self._tok[: n * self.T]
        dev.copy_(self._pinned_tok[: n * self.T], non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        return dev.view(n, self.T)

    # ---- device-resident corpus: only song indices cross PCIe ---------------------------
    def _stage_indexed(self, corpus: torch.Tensor, row_ids: np.ndarray) -> torch.Tensor:
        """corpus int32 [n_songs, T] on the device; row_ids [n] -> the step's token batch [n, T], gathered on the device."""
        ids = np.ascontiguousarray(row_ids, dtype=np.int32).reshape(-1)
        n = int(ids.size)
        if n > self.max_seqs:
            raise FsmgError(f"{n} sequences > engine capacity {self.max_seqs}")
        assert corpus.dtype == torch.int32 and corpus.is_cuda and corpus.is_contiguous() and corpus.shape[1] == self.T
        if n and (ids.min() < 0 or ids.max() >= corpus.shape[0]):
            raise FsmgError("song index outside the corpus")
        key = (corpus.data_ptr(), tuple(corpus.shape))
        if key not in self._checked_corpora:      # one device reduction per corpus, at its first use
            lo, hi = int(corpus.min()), int(corpus.max())
            if lo < 0 or hi > self.V:
                raise FsmgError(f"corpus token ids span [{lo}, {hi}], outside [0, {self.V}] (input_size = {self.V})")
            self._checked_corpora.add(key)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        if self._ids is None:
            self._ids = torch.empty(self.max_seqs, dtype=torch.int32, device=self.device)
        self._pinned_tok[:n].copy_(torch.from_numpy(ids))
        self._ids[:n].copy_(self._pinned_tok[:n], non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        dev = self._tok[: n * self.T]
        _lib.check(self.lib.fsmg_gather_token_rows(corpus.data_ptr(), int(corpus.shape[0]), self.T, self._ids.data_ptr(), n,
                                                   dev.data_ptr(), self._stream()))
        return dev.view(n, self.T)

    def train_indexed(self, corpus: torch.Tensor, row_ids: np.ndarray, global_tokens: Optional[int] = None) -> float:
        self.train_step_device(self._stage_indexed(corpus, row_ids), global_tokens)
        return self._read_loss()

    def eval_indexed(self, corpus: torch.Tensor, row_ids: np.ndarray) -> float:
        dev = self._stage_indexed(corpus, row_ids)
        _, s = self.forward_nll(dev)
        self._pinned_scal[:1].copy_(s, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return float(self._pinned_scal[0]) / (float(dev.numel()) + 1e-12)

    def _stage_rows(self, blocks) -> torch.Tensor:
        """Like _stage for a LIST of [rows, T] arrays (support / query blocks of the step's episodes): every block is copied
        once, straight into the pinned staging buffer — no intermediate concatenated array on the host."""
        n = sum(int(b.shape[0]) for b in blocks)
        if n > self.max_seqs:
            raise FsmgError(f"{n} sequences > engine capacity {self.max_seqs}")
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        if self._pinned_np is None:
            self._pinned_np = self._pinned_tok.numpy()      # shares the pinned memory
        dst = self._pinned_np[: n * self.T].reshape(n, self.T)
        r = 0
        for b in blocks:
            k = int(b.shape[0])
            if b.shape[1] != self.T:
                raise FsmgError(f"token rows of length {b.shape[1]} != max_len {self.T}")
            dst[r:r + k] = b            # numpy casts to int32 on assignment
            r += k
        self._check_ids(dst)
        dev = self._tok[: n * self.T]
        dev.copy_(self._pinned_tok[: n * self.T], non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        return dev.view(n, self.T)

    def train_host_rows(self, blocks, global_tokens: Optional[int] = None) -> float:
        self.train_step_device(self._stage_rows(blocks), global_tokens)
        return self._read_loss()

    def train_host(self, tokens: np.ndarray, global_tokens: Optional[int] = None) -> float:
        self.train_step_device(self._stage(tokens), global_tokens)
        return self._read_loss()

    def eval_host(self, tokens: np.ndarray, return_nll: bool = False):
        dev = self._stage(tokens)
        nll, s = self.forward_nll(dev)
        n_tok = dev.numel()
        self._pinned_scal[:1].copy_(s, non_blocking=True)
        out = nll.cpu().numpy().copy() if return_nll else None
        torch.cuda.current_stream(self.device).synchronize()
        mean = float(self._pinned_scal[0]) / (float(n_tok) + 1e-12)
        return (mean, out) if return_nll else mean

    def sample_host(self, n_songs: int, n_tokens: int) -> np.ndarray:
        return self.sample_greedy_device(n_songs, n_tokens).cpu().numpy()

    def set_profile(self, enable: bool) -> None:
        _lib.check(self.lib.fsmg_set_profile(self.h, int(enable)))

    def read_profile(self) -> Dict[str, dict]:
        """Per-phase device milliseconds (CUDA events on the engine's stream) since the last read."""
        n = _lib.FSMG_PROF_PHASES
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        _lib.check(self.lib.fsmg_read_profile(self.h, ms, cnt, self._stream()))
        return {self.lib.fsmg_profile_phase_name(i).decode(): dict(ms=float(ms[i]), brackets=int(cnt[i])) for i in range(n)}

    def last_launch_count(self) -> int:
        return int(self.lib.fsmg_last_launch_count(self.h))

    def close(self) -> None:
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.fsmg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
