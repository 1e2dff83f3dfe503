"""CPU ORACLE (test infrastructure, NOT product code) for the episodic-LSTM hot path.

Restates, in plain NumPy, the arithmetic of the reference's
``models.lstm_baseline.LSTMBaseline`` (reference ``src/models/lstm_baseline.py:18-156``)
together with the TensorFlow-1.x library semantics it relies on.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product path (``fsmg`` + ``libfsmg.so``) never does.

PARITY PARTLY PINNED.  Pinned to the reference itself (bit-exact, ``tests/test_reference_pin.py``
running the unmodified reference code + the committed outputs ``tests/golden/reference_*``): the
input/target shift of ``shift_inputs`` / ``episode_train_tokens`` (reference
``models/base_model.py:57-86``) that every NLL comparison feeds.  UNPINNED: the graph arithmetic
itself — it lives in TensorFlow 1.x (version unpinned, absent from ``requirements.txt``; not
installable here), and the reference's only numeric test (``src/train/test_seed.py:45-65``) needs
the real datasets and TF's initialiser RNG stream, so no golden value of it can be evaluated in
this container.  For that part the oracle is anchored on an independent formulation
(``oracle/torch_ref.py``: ``torch.nn.LSTM`` + autograd), finite differences and the committed
fixtures of ``tests/golden/make_golden.py``.

Every function cites the reference lines it follows.  ``dtype`` selects the arithmetic
(float64 = ground truth, float32 = "the reference's CPU path").
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------
# Parameters (reference lstm_baseline.py:39-40,44-49,60-62; tf_model.py:91-92)
# ----------------------------------------------------------------------------------------


def param_names(n_layers: int, scope: str = "lstm_baseline") -> List[str]:
    """TF variable names in creation order (the order ``get_vars()`` returns them,
    reference tf_model.py:99-104): embedding, per-layer kernel/bias, softmax_w, softmax_b."""
    names = [f"{scope}/embedding"]
    for layer in range(n_layers):
        base = f"{scope}/rnn/multi_rnn_cell/cell_{layer}/basic_lstm_cell"
        names += [f"{base}/kernel", f"{base}/bias"]
    names += [f"{scope}/softmax_w", f"{scope}/softmax_b"]
    return names


def param_shapes(config: dict) -> Dict[str, Tuple[int, ...]]:
    """Shapes per reference lstm_baseline.py:21-27,39-40,60-62 (V' = input_size + 1)."""
    scope = config.get("name", "lstm_baseline")
    vp = int(config["input_size"]) + 1
    e = int(config["embedding_size"])
    h = int(config["hidden_size"])
    n_layers = int(config.get("n_layers", 1))
    shapes: Dict[str, Tuple[int, ...]] = {f"{scope}/embedding": (vp, e)}
    for layer in range(n_layers):
        base = f"{scope}/rnn/multi_rnn_cell/cell_{layer}/basic_lstm_cell"
        fan = e if layer == 0 else h
        shapes[f"{base}/kernel"] = (fan + h, 4 * h)
        shapes[f"{base}/bias"] = (4 * h,)
    shapes[f"{scope}/softmax_w"] = (h, vp)
    shapes[f"{scope}/softmax_b"] = (vp,)
    return shapes


def glorot_init(config: dict, seed: int = 1234, dtype=np.float32) -> Dict[str, np.ndarray]:
    """[TF-lib] ``get_variable`` default initialiser = glorot_uniform:
    U(-l, l), l = sqrt(6/(fan_in+fan_out)); 1-D shapes use fan_in = fan_out = len, so
    softmax_b is NOT zero.  LSTM bias is zeros (BasicLSTMCell).  TF's own RNG stream is
    irreproducible, so weights are drawn from ``RandomState(seed)`` and *injected* into
    both the oracle and the engine."""
    rng = np.random.RandomState(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(config).items():
        if name.endswith("/bias"):
            out[name] = np.zeros(shape, dtype=dtype)
            continue
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        else:
            fan_in, fan_out = shape
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        out[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
    return out


# ----------------------------------------------------------------------------------------
# Inputs / targets (reference base_model.py:57-86, lstm_baseline.py:91-96,117-118)
# ----------------------------------------------------------------------------------------


def shift_inputs(tokens: np.ndarray, start_word: int) -> Tuple[np.ndarray, np.ndarray]:
    """tokens [..., T] -> X = [start, tok[:-1]], Y = tok  (base_model.py:63-86 with a
    start word; leading dims flattened like flatten_first_two_dims, :57-60)."""
    tok = np.asarray(tokens).reshape(-1, np.shape(tokens)[-1])
    x = np.empty_like(tok)
    x[:, 0] = start_word
    x[:, 1:] = tok[:, :-1]
    return x, tok.copy()


def episode_train_tokens(support: np.ndarray, query: np.ndarray) -> np.ndarray:
    """train(): support rows first, then query rows (lstm_baseline.py:91-96)."""
    t = support.shape[-1]
    return np.concatenate([support.reshape(-1, t), query.reshape(-1, t)], axis=0)


# ----------------------------------------------------------------------------------------
# Forward (reference lstm_baseline.py:38-75)
# ----------------------------------------------------------------------------------------


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _layer_params(params: Dict[str, np.ndarray], scope: str, n_layers: int):
    ks, bs = [], []
    for layer in range(n_layers):
        base = f"{scope}/rnn/multi_rnn_cell/cell_{layer}/basic_lstm_cell"
        ks.append(params[f"{base}/kernel"])
        bs.append(params[f"{base}/bias"])
    return ks, bs


def _scope_of(params: Dict[str, np.ndarray]) -> str:
    for k in params:
        if k.endswith("/embedding"):
            return k[: -len("/embedding")]
    raise KeyError("no embedding in params")


def _n_layers_of(params: Dict[str, np.ndarray]) -> int:
    return sum(1 for k in params if k.endswith("/kernel"))


def forward(params: Dict[str, np.ndarray], x: np.ndarray, y: np.ndarray, dtype=np.float64,
            keep_cache: bool = False):
    """Per-token NLL [N,T] and mean loss.

    embedding_lookup + unstack (lstm_baseline.py:39-42); BasicLSTMCell(forget_bias=1)
    gates i,j,f,o (:44-49, [TF-lib] A.2); zero initial state, full-length static_rnn
    (:50-55, A.3); xw_plus_b (:57-67); sequence_loss with all-ones weights = sum(nll) /
    (N*T + 1e-12), padding id 0 is scored (:70-75, A.5).
    """
    scope = _scope_of(params)
    n_layers = _n_layers_of(params)
    p = {k: v.astype(dtype) for k, v in params.items()}
    emb = p[f"{scope}/embedding"]
    ks, bs = _layer_params(p, scope, n_layers)
    sw, sb = p[f"{scope}/softmax_w"], p[f"{scope}/softmax_b"]
    n, t_steps = x.shape
    hsz = sw.shape[0]
    inp = emb[x]  # [N,T,E]
    cache = {"x": x, "y": y, "layers": []}
    for layer in range(n_layers):
        k, b = ks[layer], bs[layer]
        c = np.zeros((n, hsz), dtype)
        h = np.zeros((n, hsz), dtype)
        hs = np.empty((n, t_steps, hsz), dtype)
        lc = {"inp": inp, "i": [], "j": [], "f": [], "o": [], "c": [], "c_prev": [], "h_prev": []}
        for t in range(t_steps):
            g = np.concatenate([inp[:, t, :], h], axis=1) @ k + b
            gi, gj, gf, go = np.split(g, 4, axis=1)
            i_, j_, f_, o_ = _sigmoid(gi), np.tanh(gj), _sigmoid(gf + dtype(1.0)), _sigmoid(go)
            if keep_cache:
                lc["c_prev"].append(c)
                lc["h_prev"].append(h)
            c = c * f_ + i_ * j_
            h = np.tanh(c) * o_
            hs[:, t, :] = h
            if keep_cache:
                for nm, val in (("i", i_), ("j", j_), ("f", f_), ("o", o_), ("c", c)):
                    lc[nm].append(val)
        lc["hs"] = hs
        cache["layers"].append(lc)
        inp = hs
    logits = inp.reshape(n * t_steps, hsz) @ sw + sb  # row order n*T+t (sequence-major)
    m = logits.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True)))[:, 0]
    tgt = logits[np.arange(n * t_steps), y.reshape(-1)]
    nll = (lse - tgt).reshape(n, t_steps)
    loss = nll.sum() / (dtype(n * t_steps) + dtype(1e-12))
    if keep_cache:
        cache.update(logits=logits, lse=lse, final_state=(c, h))
        return nll, loss, cache
    return nll, loss


def per_token_nll(params, tokens: np.ndarray, start_word: int, dtype=np.float64) -> np.ndarray:
    x, y = shift_inputs(tokens, start_word)
    return forward(params, x, y, dtype)[0]


def eval_episode(params, query: np.ndarray, start_word: int, dtype=np.float64) -> float:
    """LSTMBaseline.eval: query set only (lstm_baseline.py:115-133)."""
    x, y = shift_inputs(query, start_word)
    return float(forward(params, x, y, dtype)[1])


# ----------------------------------------------------------------------------------------
# Backward (what tf.gradients computes, lstm_baseline.py:83-84) — analytic BPTT
# ----------------------------------------------------------------------------------------


def backward(params, cache, dtype=np.float64, loss_denominator: float | None = None):
    """Dense gradients of the mean loss + the per-occurrence embedding-gradient square
    norm needed by TF's clip_by_global_norm on IndexedSlices.

    ``loss_denominator`` overrides N*T (used to emulate a data-parallel shard whose
    loss is normalised by the *global* token count)."""
    scope = _scope_of(params)
    n_layers = _n_layers_of(params)
    p = {k: v.astype(dtype) for k, v in params.items()}
    ks, _ = _layer_params(p, scope, n_layers)
    sw = p[f"{scope}/softmax_w"]
    x, y = cache["x"], cache["y"]
    n, t_steps = x.shape
    hsz = sw.shape[0]
    denom = dtype(loss_denominator if loss_denominator is not None else n * t_steps) + dtype(1e-12)

    logits, lse = cache["logits"], cache["lse"]
    dlogits = np.exp(logits - lse[:, None])
    dlogits[np.arange(n * t_steps), y.reshape(-1)] -= 1.0
    dlogits /= denom
    top_h = cache["layers"][-1]["hs"].reshape(n * t_steps, hsz)
    grads: Dict[str, np.ndarray] = {}
    grads[f"{scope}/softmax_w"] = top_h.T @ dlogits
    grads[f"{scope}/softmax_b"] = dlogits.sum(axis=0)
    dout = (dlogits @ sw.T).reshape(n, t_steps, hsz)  # dL/dh^{top}_t from the projection

    for layer in reversed(range(n_layers)):
        lc = cache["layers"][layer]
        k = ks[layer]
        fan = k.shape[0] - hsz
        inp = lc["inp"]
        dk = np.zeros_like(k)
        db = np.zeros(4 * hsz, dtype)
        dinp = np.empty((n, t_steps, fan), dtype)
        dh_next = np.zeros((n, hsz), dtype)
        dc_next = np.zeros((n, hsz), dtype)
        for t in reversed(range(t_steps)):
            i_, j_, f_, o_, c = lc["i"][t], lc["j"][t], lc["f"][t], lc["o"][t], lc["c"][t]
            c_prev, h_prev = lc["c_prev"][t], lc["h_prev"][t]
            dh = dout[:, t, :] + dh_next
            tc = np.tanh(c)
            do = dh * tc
            dc = dh * o_ * (1.0 - tc * tc) + dc_next
            dgi = dc * j_ * i_ * (1.0 - i_)
            dgj = dc * i_ * (1.0 - j_ * j_)
            dgf = dc * c_prev * f_ * (1.0 - f_)
            dgo = do * o_ * (1.0 - o_)
            dg = np.concatenate([dgi, dgj, dgf, dgo], axis=1)
            xin = np.concatenate([inp[:, t, :], h_prev], axis=1)
            dk += xin.T @ dg
            db += dg.sum(axis=0)
            dxin = dg @ k.T
            dinp[:, t, :] = dxin[:, :fan]
            dh_next = dxin[:, fan:]
            dc_next = dc * f_
        base = f"{scope}/rnn/multi_rnn_cell/cell_{layer}/basic_lstm_cell"
        grads[f"{base}/kernel"] = dk
        grads[f"{base}/bias"] = db
        dout = dinp

    # embedding: IndexedSlices(values=dout[N*T,E], indices=x) — densify for Adam, but the
    # clip norm is over the *un-aggregated* rows (A.6)
    emb = p[f"{scope}/embedding"]
    demb = np.zeros_like(emb)
    np.add.at(demb, x.reshape(-1), dout.reshape(n * t_steps, -1))
    grads[f"{scope}/embedding"] = demb
    occ_sqnorm = float((dout.astype(np.float64) ** 2).sum())
    return grads, occ_sqnorm


# ----------------------------------------------------------------------------------------
# Clip + Adam + LR schedule (lstm_baseline.py:77-87)
# ----------------------------------------------------------------------------------------


def global_norm(grads: Dict[str, np.ndarray], occ_sqnorm: float) -> float:
    """[TF-lib] clip_by_global_norm: sqrt(sum dense ||g||^2 + ||IndexedSlices.values||^2)."""
    total = occ_sqnorm
    for name, g in grads.items():
        if name.endswith("/embedding"):
            continue
        total += float((g.astype(np.float64) ** 2).sum())
    return math.sqrt(total)


def lr_at(lr0: float, step: int, n_decay: int, dtype=np.float32) -> float:
    """exponential_decay(lr, global_step, n_decay, 0.5, staircase=False) on the
    pre-increment step (lstm_baseline.py:77-81, A.8), fp32 arithmetic."""
    return float(dtype(lr0) * np.power(dtype(0.5), dtype(step) / dtype(n_decay)))


class TrainState:
    """Parameters + TF-Adam slots + global_step (tf_model.py:92; A.1, A.7)."""

    def __init__(self, params: Dict[str, np.ndarray], config: dict, dtype=np.float64):
        self.dtype = dtype
        self.params = {k: v.astype(dtype) for k, v in params.items()}
        self.m = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.step = 0  # global_step == number of applies so far
        self.lr = float(config.get("lr", 5e-3))
        self.n_decay = int(config.get("n_decay", 10000))
        self.max_grad_norm = float(config.get("max_grad_norm", 5))
        self.start_word = int(config["input_size"])
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8


def apply_clip_adam(state: TrainState, grads: Dict[str, np.ndarray], occ_sqnorm: float) -> float:
    """clip_by_global_norm then AdamOptimizer.apply_gradients ([TF-lib] A.6, A.7):
    scale = clip/max(norm, clip); alpha_t = lr_k*sqrt(1-b2^t)/(1-b1^t);
    theta -= alpha_t * m / (sqrt(v) + eps)  (eps OUTSIDE the bias correction);
    the embedding's sparse update equals dense Adam on the densified gradient."""
    dt = state.dtype
    norm = global_norm(grads, occ_sqnorm)
    scale = state.max_grad_norm / max(norm, state.max_grad_norm)
    lr_k = lr_at(state.lr, state.step, state.n_decay)
    t = state.step + 1
    alpha = lr_k * math.sqrt(1.0 - state.beta2 ** t) / (1.0 - state.beta1 ** t)
    for name, g in grads.items():
        g = (g * dt(scale)).astype(dt)
        state.m[name] = dt(state.beta1) * state.m[name] + dt(1.0 - state.beta1) * g
        state.v[name] = dt(state.beta2) * state.v[name] + dt(1.0 - state.beta2) * g * g
        state.params[name] = state.params[name] - dt(alpha) * state.m[name] / (np.sqrt(state.v[name]) + dt(state.eps))
    state.step += 1
    return norm


def train_step(state: TrainState, tokens: np.ndarray) -> float:
    """One ``LSTMBaseline.train`` call on already-concatenated rows [N,T]
    (lstm_baseline.py:89-113): returns the pre-update mean loss."""
    x, y = shift_inputs(tokens, state.start_word)
    _, loss, cache = forward(state.params, x, y, state.dtype, keep_cache=True)
    grads, occ = backward(state.params, cache, state.dtype)
    apply_clip_adam(state, grads, occ)
    return float(loss)


# ----------------------------------------------------------------------------------------
# Greedy sampling (reference lstm_baseline.py:135-156)
# ----------------------------------------------------------------------------------------


def sample_greedy(params, num: int, dtype=np.float64, return_margins: bool = False):
    """word0 = start id V, zero state; each step: one cell step on embedding[word],
    p = softmax(h W + b), word = first-max argmax (np.argmax, :152-153).  The support
    set is ignored (:135-136).  Output may contain id V."""
    scope = _scope_of(params)
    n_layers = _n_layers_of(params)
    p = {k: v.astype(dtype) for k, v in params.items()}
    emb = p[f"{scope}/embedding"]
    ks, bs = _layer_params(p, scope, n_layers)
    sw, sb = p[f"{scope}/softmax_w"], p[f"{scope}/softmax_b"]
    hsz = sw.shape[0]
    word = emb.shape[0] - 1
    cs = [np.zeros(hsz, dtype) for _ in range(n_layers)]
    hs = [np.zeros(hsz, dtype) for _ in range(n_layers)]
    out, margins = [], []
    for _ in range(num):
        inp = emb[word]
        for layer in range(n_layers):
            g = np.concatenate([inp, hs[layer]]) @ ks[layer] + bs[layer]
            gi, gj, gf, go = np.split(g, 4)
            cs[layer] = cs[layer] * _sigmoid(gf + dtype(1.0)) + _sigmoid(gi) * np.tanh(gj)
            hs[layer] = np.tanh(cs[layer]) * _sigmoid(go)
            inp = hs[layer]
        logits = inp @ sw + sb
        word = int(np.argmax(logits))  # softmax is monotone: argmax(prob) == argmax(logits)
        out.append(word)
        if return_margins:
            top2 = np.partition(logits, -2)[-2:]
            margins.append(float(top2[1] - top2[0]))
    return (out, margins) if return_margins else out


def greedy_deficits(params, tokens, dtype=np.float64) -> np.ndarray:
    """Teacher-forced check of a greedy decode produced elsewhere (the GPU): feed the GIVEN tokens back as inputs
    (reference lstm_baseline.py:135-156 with `word` replaced by tokens[i]) and return, per step, how far the given token's
    logit lies below the step's maximum: 0 where the token IS the argmax.  Unlike a token-by-token comparison with
    `sample_greedy`, this keeps checking after a numerical tie (where two correct decoders legitimately diverge): every
    step is judged against the oracle's own logits for the sequence actually generated."""
    scope = _scope_of(params)
    n_layers = _n_layers_of(params)
    p = {k: v.astype(dtype) for k, v in params.items()}
    emb = p[f"{scope}/embedding"]
    ks, bs = _layer_params(p, scope, n_layers)
    sw, sb = p[f"{scope}/softmax_w"], p[f"{scope}/softmax_b"]
    hsz = sw.shape[0]
    word = emb.shape[0] - 1
    cs = [np.zeros(hsz, dtype) for _ in range(n_layers)]
    hs = [np.zeros(hsz, dtype) for _ in range(n_layers)]
    out = np.zeros(len(tokens), dtype)
    for i, given in enumerate(tokens):
        inp = emb[word]
        for layer in range(n_layers):
            g = np.concatenate([inp, hs[layer]]) @ ks[layer] + bs[layer]
            gi, gj, gf, go = np.split(g, 4)
            cs[layer] = cs[layer] * _sigmoid(gf + dtype(1.0)) + _sigmoid(gi) * np.tanh(gj)
            hs[layer] = np.tanh(cs[layer]) * _sigmoid(go)
            inp = hs[layer]
        logits = inp @ sw + sb
        out[i] = logits.max() - logits[int(given)]
        word = int(given)
    return out


# ----------------------------------------------------------------------------------------
# Synthetic episodes (SURVEY §8 d) — shared by tests and bench
# ----------------------------------------------------------------------------------------


def synthetic_tokens(rng: np.random.RandomState, shape, vocab: int, kind: str = "zipf") -> np.ndarray:
    """Token ids in [0, vocab): Zipf(s=1.0) truncated (lyrics-like) or uniform (MIDI-like)."""
    if kind == "zipf":
        ranks = np.arange(1, vocab + 1, dtype=np.float64)
        pmf = 1.0 / ranks
        cdf = np.cumsum(pmf / pmf.sum())
        u = rng.random_sample(size=shape)
        return np.minimum(np.searchsorted(cdf, u), vocab - 1).astype(np.int32)
    if kind == "uniform":
        return rng.randint(0, vocab, size=shape).astype(np.int32)
    raise ValueError(kind)


def synthetic_episode(rng, batch_size: int, support_size: int, query_size: int, max_len: int,
                      vocab: int, kind: str = "zipf", pad_fraction: float = 0.0):
    """support [B,S,T], query [B,Q,T] int32 like EpisodeSampler.get_episode
    (reference data/episode.py:62-74); optional zero-padding tail (base_loader.py:52-64)."""
    sup = synthetic_tokens(rng, (batch_size, support_size, max_len), vocab, kind)
    qry = synthetic_tokens(rng, (batch_size, query_size, max_len), vocab, kind)
    if pad_fraction > 0:
        for arr in (sup, qry):
            flat = arr.reshape(-1, max_len)
            for r in range(flat.shape[0]):
                if rng.random_sample() < pad_fraction:
                    flat[r, rng.randint(1, max_len):] = 0
    return sup, qry
