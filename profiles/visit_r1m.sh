FSMG_SAMPLE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r1m_sample_launches.csv python bench.py --mode sample --steps 1 --warmup 1 > gpurun_out/r1m_sample.log 2>&1
tail -2 gpurun_out/r1m_sample.log
