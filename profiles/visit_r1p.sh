timeout 900 python -m pytest tests -m gpu -x -q -k "sampl or greedy or plugin" 2>&1 | tail -2
for u in 16 1; do echo "unroll $u"; FSMG_SAMPLE_UNROLL=$u timeout 300 python bench.py --mode sample --steps 5 --warmup 2 | cut -c80-200; done
