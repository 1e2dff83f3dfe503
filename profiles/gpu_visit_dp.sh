#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): data-parallel parity test + the bench line at N ranks (dp_check inside), overlap on and off.
# usage: bash profiles/gpu_visit_dp.sh <tag> <N>
tag=${1:-r2dp}
n=${2:-2}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dp.py -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
for ov in ${3:-0 1}; do
FSMG_AR_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline $( [ $ov = 1 ] && echo --no-extra-configs ) > $out/${tag}_bench_ov$ov.json 2> $out/${tag}_bench_ov$ov.err
echo "bench overlap=$ov rc=$?"; tail -3 $out/${tag}_bench_ov$ov.err
python profiles/phases.py < $out/${tag}_bench_ov$ov.json
python - <<PY
import json
d=json.loads([l for l in open("$out/${tag}_bench_ov$ov.json") if l.startswith("{")][-1])
print("dp_check:", json.dumps(d.get("dp_check")))
PY
done
