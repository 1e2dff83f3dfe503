"""Does train() return before the step has finished (FSMG_EARLY_LOSS)?  Per-call host wall time, the time a final synchronize still
has to wait, and the host-side staging time, for the configs[1] batch through the plugin class."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from train.train import load_model_from_config  # noqa: E402

w = bench.WORKLOADS["lyrics5shot_v10k_t128_h512"]
cfg = bench.model_config(w)
cfg["episodes_per_step"] = w["episodes"]
model = load_model_from_config(cfg)
model.recover_or_init("")
batches = bench.synthetic_batches(w, 4, 1234)


class Ep:
    def __init__(self, s, q):
        self.support, self.query = s, q


hb = [[Ep(s, q) for s, q in b] for b in batches]
for i in range(4):
    model.train(hb[i % 4])
torch.cuda.synchronize()
calls = []
t_all = time.perf_counter()
for i in range(12):
    t0 = time.perf_counter()
    model.train(hb[i % 4])
    calls.append((time.perf_counter() - t0) * 1e3)
t_loop = (time.perf_counter() - t_all) * 1e3
t0 = time.perf_counter()
torch.cuda.synchronize()
t_tail = (time.perf_counter() - t0) * 1e3
eng = model.engine
t0 = time.perf_counter()
for i in range(12):
    blocks = []
    for ep in hb[i % 4]:
        blocks.append(ep.support.reshape(-1, 128))
        blocks.append(ep.query.reshape(-1, 128))
    eng._stage_rows(blocks)
torch.cuda.synchronize()
t_stage = (time.perf_counter() - t0) * 1e3 / 12
print(f"early_loss={eng._early_loss} per-call ms {np.round(calls, 2).tolist()}")
print(f"loop {t_loop:.2f} ms for 12 steps = {t_loop / 12:.3f} ms/step; final synchronize waited {t_tail:.2f} ms; host staging alone {t_stage:.3f} ms/step")
