#!/bin/bash
# One GPU-box visit (round 2): parity tests, both bench arms (the product arm carries the cfg3 / cfg5 sub-records), launch lists.
# usage (under gpurun): bash profiles/gpu_visit.sh <tag> [quick]
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python profiles/phases.py < $out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench_ref.json
if [ "$2" != "quick" ]; then
# launch list of the decode loop (cold-cache, serialised: shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $out/${tag}_sample_launches.csv \
    python bench.py --mode sample --steps 1 --warmup 1 > $out/${tag}_sample_launches.log 2>&1
python profiles/summarize_launches.py $out/${tag}_sample_launches.csv > $out/${tag}_sample_launches.md 2>&1; head -20 $out/${tag}_sample_launches.md
FSMG_TRACE=1 timeout 300 python profiles/profile_step.py 1 > $out/${tag}_trace.log 2>&1; tail -20 $out/${tag}_trace.log
fi
ls -la $out | grep ${tag}_
