for u in 16 1 4 64; do echo "unroll $u"; FSMG_SAMPLE_UNROLL=$u timeout 300 python bench.py --mode sample --steps 5 --warmup 2 | cut -c80-200; done
FSMG_SAMPLE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 16 --csv --log-file gpurun_out/r1o_sample_launches.csv python bench.py --mode sample --steps 1 --warmup 1 > /dev/null 2>&1
