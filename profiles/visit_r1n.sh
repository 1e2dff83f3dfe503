timeout 900 python -m pytest tests -m gpu -x -q -k "sampl or greedy or plugin" 2>&1 | tail -4
timeout 300 python bench.py --mode sample --steps 5 --warmup 2 | tee gpurun_out/r1n_sample.json | cut -c1-400
