#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=$1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_abi_host.py -q -k "sampl or greedy or host_abi" > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py --mode sample --steps 3 --warmup 1 > $out/${tag}_sample.json 2> $out/${tag}_sample.err; python -c "
import json;d=json.loads([l for l in open('$out/${tag}_sample.json') if l.startswith('{')][-1]);print('sample Mtok/s %.2f  us/step %.1f launches %d'%(d['value']/1e6,d['us_per_token_step'],d['gpu_launches']))"
