#!/bin/bash
# GPU-box visit for the fused softmax gradient: transform-GEMM unit tests, engine A/B parity, then bench A/B.
# usage (under gpurun): bash profiles/gpu_visit_xf.sh <tag> [tests-only]
tag=${1:-r3x}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests/test_gpu_xf.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -25 $out/${tag}_pytest.log
if [ "$2" != "tests-only" ]; then
bash profiles/ab.sh $tag "FSMG_FUSED_SG=0" "FSMG_FUSED_SG=1"
fi
