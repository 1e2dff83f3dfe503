timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1u_bench.json 2> gpurun_out/r1u_bench.err; tail -3 gpurun_out/r1u_bench.err
python profiles/phases.py < gpurun_out/r1u_bench.json
python -c "
import json;d=json.loads(open('gpurun_out/r1u_bench.json').read().strip().splitlines()[-1]);print(d['e2e'])"
