"""Seeded synthetic episodes for the benchmarks (SURVEY §8d "Synthetic inputs"): token ids are Zipf(s=1.0)
truncated to [0, V) (lyrics-like) or uniform over [0, V) (MIDI-event-like); an episode is
``support [B,S,T]`` + ``query [B,Q,T]`` int32 like ``EpisodeSampler.get_episode`` returns
(reference ``src/data/episode.py:62-74``).  No filesystem, no dependency on the test oracle."""
from __future__ import annotations

import numpy as np


def synthetic_tokens(rng: np.random.RandomState, shape, vocab: int, kind: str = "zipf") -> np.ndarray:
    if kind == "zipf":
        pmf = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
        cdf = np.cumsum(pmf / pmf.sum())
        return np.minimum(np.searchsorted(cdf, rng.random_sample(size=shape)), vocab - 1).astype(np.int32)
    if kind == "uniform":
        return rng.randint(0, vocab, size=shape).astype(np.int32)
    raise ValueError(kind)


def synthetic_episode(rng: np.random.RandomState, batch_size: int, support_size: int, query_size: int, max_len: int,
                      vocab: int, kind: str = "zipf"):
    sup = synthetic_tokens(rng, (batch_size, support_size, max_len), vocab, kind)
    qry = synthetic_tokens(rng, (batch_size, query_size, max_len), vocab, kind)
    return sup, qry
