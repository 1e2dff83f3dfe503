"""Per-phase device times of the scoring pass (fsmg_forward_nll: no logits store, no backward) at BASELINE configs[1]."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fsmg.engine import Engine  # noqa: E402
from data import synthetic as S  # noqa: E402
w = bench.WORKLOADS["lyrics5shot_v10k_t128_h512"]
cfg = bench.model_config(w)
n = bench.SEQS_PER_EPISODE * w["episodes"]
eng = Engine(cfg, max_seqs=n, device="cuda:0")
eng.init_params(1234)
tok = torch.from_numpy(S.synthetic_tokens(np.random.RandomState(0), (n, w["max_len"]), w["input_size"], "zipf")).cuda()
for _ in range(3):
    eng.forward_nll(tok)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.forward_nll(tok)
e1.record(); torch.cuda.synchronize()
print("eval ms/pass %.3f  (%.1f M tok/s)" % (e0.elapsed_time(e1) / 10, tok.numel() / (e0.elapsed_time(e1) / 10) / 1e3))
eng.set_profile(True); eng.read_profile()
for _ in range(2):
    eng.forward_nll(tok)
print({k: round(v["ms"] / 2, 3) for k, v in eng.read_profile().items() if v["ms"] > 0})
