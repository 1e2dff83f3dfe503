"""Host-side breakdown of one end-to-end training step (numpy episodes in -> loss out): where the CPU time of
LSTMBaseline.train goes when every step synchronises on its loss.  Prints medians over the timed steps."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

w = bench.WORKLOADS["lyrics5shot_v10k_t128_h512"]
cfg = bench.model_config(w)
cfg["episodes_per_step"] = w["episodes"]
from train.train import load_model_from_config  # noqa: E402  (bench put src/ on sys.path)

model = load_model_from_config(cfg)
model.recover_or_init("")
eng = model.engine
batches = bench.synthetic_batches(w, 2, 1234)


class Ep:
    def __init__(self, s, q):
        self.support, self.query = s, q


hb = [[Ep(s, q) for s, q in b] for b in batches]
rows = {k: [] for k in ("concat", "stage", "fwd_bwd_enqueue", "update_enqueue", "loss_sync", "total")}
for i in range(14):
    t0 = time.perf_counter()
    tok = model._train_tokens(hb[i % 2])
    t1 = time.perf_counter()
    dev = eng._stage(tok)
    t2 = time.perf_counter()
    n = int(dev.shape[0])
    eng.forward_backward(dev, n * eng.T)
    t3 = time.perf_counter()
    from fsmg import _lib
    _lib.check(eng.lib.fsmg_apply_update(eng.h, eng.global_step, 0, eng._stream()))
    eng.global_step += 1
    t4 = time.perf_counter()
    loss = float((eng.grads[eng.n_params: eng.n_params + 1] / (n * eng.T + 1e-12)).cpu())
    t5 = time.perf_counter()
    if i >= 4:
        for k, v in zip(rows, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
            rows[k].append(v * 1e3)
print("e2e host breakdown (ms, median of 10): " + "  ".join("%s=%.3f" % (k, float(np.median(v))) for k, v in rows.items()), "loss", loss)
