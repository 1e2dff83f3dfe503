"""Worker for tests/test_gpu_dp.py (launched by torch.distributed.run, one rank per GPU):
k-GPU data-parallel step on N/k sequences per rank == 1-GPU step on all N sequences."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "few-shot-music-generation_b200"
for p in (str(ROOT), str(PKG), str(PKG / "src")):
    sys.path.insert(0, p)

from oracle import lstm_oracle as O  # noqa: E402  (test-side checker + synthetic inputs)
from fsmg.engine import Engine  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = dict(name="lstm_baseline", input_size=300, embedding_size=64, hidden_size=64, n_layers=1, max_len=8,
               lr=5e-3, n_decay=10000, max_grad_norm=0.5)
    per_rank = 24
    params = O.glorot_init(cfg, 5)
    tok = O.synthetic_tokens(np.random.RandomState(11), (per_rank * world, 8), 300, "zipf")
    eng = Engine(cfg, max_seqs=per_rank, device=f"cuda:{local}")
    eng.load_params(params)
    shard = torch.from_numpy(tok[rank * per_rank:(rank + 1) * per_rank]).cuda()
    losses = [float(eng.train_step_device(shard)) for _ in range(3)]
    mine = eng.export("params")
    # replicas stay bit-identical: same reduced gradients, same update
    flat = eng.params.clone()
    dist.broadcast(flat, src=0)
    assert torch.equal(flat, eng.params), "replicas diverged"
    if rank == 0:
        state = O.TrainState(params, cfg, np.float64)
        ref_losses = [O.train_step(state, tok) for _ in range(3)]
        np.testing.assert_allclose(losses, ref_losses, rtol=1e-3)
        for k, v in mine.items():
            # norm-wise (Adam turns a near-zero gradient whose sign differs by rounding into a full-size step for that element)
            upd = np.linalg.norm(state.params[k] - params[k]) + 1e-12
            assert np.linalg.norm(v - state.params[k]) < 0.05 * upd, (k, np.linalg.norm(v - state.params[k]), upd)
        print("DP_OK world=%d losses=%s" % (world, np.round(losses, 5)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
