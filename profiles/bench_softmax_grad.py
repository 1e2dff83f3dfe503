"""Micro-benchmark of the in-place softmax-gradient pass (the one purely HBM-bound kernel of the training step) through the
C-ABI debug entry: every kernel variant over a BASELINE configs[1]-sized logits chunk (18 432 x 10 016 fp16 = 369 MB, > L2),
CUDA events on the launching stream, algorithmic bytes = 2 x rows x ld x 2 B (one read + one write of the chunk)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402,F401  (path setup)
from fsmg import _lib  # noqa: E402

lib = _lib.load()
rows, v1 = 18432, 10001
ld = 10016
dev = "cuda:0"
torch.manual_seed(0)
src = (torch.randn(rows, ld, device=dev) * 2).half()
lse = torch.logsumexp(src[:, :v1].float(), dim=1).contiguous()
y = torch.randint(0, v1, (rows,), device=dev, dtype=torch.int32)
db = torch.zeros(v1, device=dev)
bufs = [src.clone() for _ in range(2)]          # rotate 2 x 369 MB so no launch finds its input in L2
variants = [(0, 16, 6), (0, 32, 6), (0, 64, 6), (0, 128, 6), (1, 16, 6), (1, 32, 6), (1, 32, 4), (2, 1, 6), (2, 2, 6), (2, 4, 6),
            (2, 2, 4), (2, 4, 4), (2, 2, 3), (2, 4, 3), (2, 4, 2)]
if len(sys.argv) > 1:
    variants = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
s = torch.cuda.current_stream().cuda_stream
nbytes = 2.0 * rows * ld * 2
for mode, param, waves in variants:
    def go(i):
        _lib.check(lib.fsmg_debug_softmax_grad(rows, v1, ld, bufs[i % 2].data_ptr(), lse.data_ptr(), y.data_ptr(), 1.0, db.data_ptr(),
                                               mode, param, waves, s))
    for i in range(4):
        go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        go(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"mode {mode} param {param:3d} waves {waves}: {us:7.1f} us  {nbytes / us / 1e3:7.1f} GB/s", flush=True)
