#!/bin/bash
# recurrent-kernel experiments: parity tests first (under a hard timeout: a protocol bug would hang), then trace + A/B bench per variant
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_parity.py -q -x -k "not entry_point" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
bash profiles/ab_lstm.sh $tag "$@"
