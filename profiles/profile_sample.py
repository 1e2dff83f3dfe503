"""A short greedy decode of BASELINE configs[4] dims (256 songs, E=H=1024, V=4708): the command ncu wraps for the decode-step kernels."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fsmg.engine import Engine  # noqa: E402

n_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = bench.model_config(bench.WORKLOADS["midi5shot_v4708_t256_h1024"])
eng = Engine(cfg, max_seqs=256, device="cuda:0")
eng.init_params(1234)
out = eng.sample_greedy_device(256, n_tokens)
torch.cuda.synchronize()
print("ok", out[0, :8].tolist())
