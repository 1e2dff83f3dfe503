"""Generates the committed golden fixtures from the fp64 oracle (oracle/lstm_oracle.py).

    python tests/golden/make_golden.py

The reference itself (TensorFlow 1.x) cannot run in this container, so these vectors pin the
ORACLE (and through it the CUDA path), not TensorFlow: "parity unpinned" per SURVEY §8c.
Weights are re-derived from the seed (RandomState is stable across NumPy versions); a checksum of
the parameters is stored so drift is detected.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import lstm_oracle as O  # noqa: E402

CASES = {
    # name: (config, n_episodes, token kind, pad_fraction, n_updates, n_sample)
    "tiny_l1": (dict(input_size=50, embedding_size=12, hidden_size=10, n_layers=1, max_len=6), 1, "uniform", 0.0, 10, 6),
    "tiny_l2": (dict(input_size=50, embedding_size=12, hidden_size=10, n_layers=2, max_len=6), 1, "uniform", 0.3, 10, 6),
    "odd_dims": (dict(input_size=333, embedding_size=50, hidden_size=36, n_layers=1, max_len=9), 2, "zipf", 0.2, 10, 9),
    # BASELINE.json configs[0]: 1 episode, seq_len=32, hidden=128, E=250 (reference default), lyrics-like V
    "cfg1_cpu_ref": (dict(input_size=10000, embedding_size=250, hidden_size=128, n_layers=1, max_len=32), 1, "zipf", 0.0, 10, 32),
}


def build_case(name):
    cfg, n_ep, kind, pad, n_upd, n_samp = CASES[name]
    cfg = dict(cfg, name="lstm_baseline", lr=5e-3, n_decay=10000, max_grad_norm=5)
    params = O.glorot_init(cfg, 1234)
    rng = np.random.RandomState(4321)
    batches = []
    for _ in range(n_upd):
        rows = []
        for _ in range(n_ep):
            sup, qry = O.synthetic_episode(rng, 5, 5, 4, cfg["max_len"], cfg["input_size"], kind, pad)
            rows.append(O.episode_train_tokens(sup, qry))
        batches.append(np.concatenate(rows))
    tokens = np.stack(batches)  # [n_upd, N, T]
    nll0 = O.per_token_nll(params, tokens[0], cfg["input_size"], np.float64)
    state = O.TrainState(params, cfg, np.float64)
    losses = [O.train_step(state, tokens[i]) for i in range(n_upd)]
    nll_after = O.per_token_nll(state.params, tokens[0], cfg["input_size"], np.float64)
    sample, margins = O.sample_greedy(params, n_samp, np.float64, True)
    checksum = float(sum(np.float64(v).sum() for v in params.values()))
    return cfg, dict(tokens=tokens.astype(np.int32), nll_initial=nll0, losses=np.asarray(losses),
                     nll_after_training=nll_after, sample=np.asarray(sample, np.int32),
                     sample_margins=np.asarray(margins), param_checksum=np.asarray(checksum))


if __name__ == "__main__":
    out = Path(__file__).resolve().parent
    for name in CASES:
        cfg, blob = build_case(name)
        np.savez_compressed(out / f"{name}.npz", **blob)
        print(name, "losses", blob["losses"][[0, -1]], "sample", blob["sample"][:8], "min margin", blob["sample_margins"].min())
