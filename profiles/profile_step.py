"""Runs a few device-resident training steps of BASELINE configs[1] (no CPU baseline, no e2e arm):
the short command that ncu wraps for the per-kernel captures under profiles/."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "lyrics5shot_v10k_t128_h512"]
cfg = bench.model_config(w)
from fsmg.engine import Engine  # noqa: E402

n = bench.SEQS_PER_EPISODE * w["episodes"]
eng = Engine(cfg, max_seqs=n, device="cuda:0")
eng.init_params(1234)
rng = np.random.RandomState(0)
from data import synthetic as O  # noqa: E402

tok = torch.from_numpy(O.synthetic_tokens(rng, (n, w["max_len"]), w["input_size"], w["kind"])).cuda()
for _ in range(steps):
    eng.train_step_device(tok)
torch.cuda.synchronize()
print("ok", float(eng.grads[eng.n_params]) / tok.numel())
