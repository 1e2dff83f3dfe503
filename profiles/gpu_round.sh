#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, the ncu launch list and the --set full captures.
# usage (under gpurun): bash profiles/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python profiles/phases.py < $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
timeout 300 python bench.py --mode sample --steps 3 --warmup 2 > $out/${tag}_bench_sample.json 2>> $out/${tag}_bench.err
# launch list of the bench command's hot loop (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python profiles/profile_step.py 2 > $out/${tag}_launches.log 2>&1
# --set full of the GEMM-core instantiations (first step is graph capture + warm; take launches of the second step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 12 -c 10 -f -o $out/${tag}_gemm \
    python profiles/profile_step.py 2 > $out/${tag}_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lstm_.*_kernel|softmax_grad' -s 4 -c 4 -f -o $out/${tag}_lstm \
    python profiles/profile_step.py 2 > $out/${tag}_ncu_lstm.log 2>&1
ls -la $out | tail -20
