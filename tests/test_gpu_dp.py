"""Multi-GPU data-parallel parity (needs >= 2 GPUs: `gpurun --gpus 2`): one process per GPU over
NCCL; the sharded step must reproduce the oracle's single-process step on the union batch."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("overlap", ["0", "1"], ids=["single_allreduce", "stage_event_overlap"])
def test_two_gpu_data_parallel_step_matches_single_process(built_lib, overlap):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 8)
    worker = Path(__file__).with_name("dp_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), str(worker)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, FSMG_AR_OVERLAP=overlap))
    (Path(__file__).resolve().parents[1] / "gpurun_out").mkdir(exist_ok=True)
    (Path(__file__).resolve().parents[1] / "gpurun_out" / "dp_worker_last.log").write_text(out.stdout[-20000:] + "\n==== stderr ====\n" + out.stderr[-20000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DP_OK" in out.stdout
