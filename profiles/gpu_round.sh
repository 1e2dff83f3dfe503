#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, the ncu launch list and the --set full captures.
# usage (under gpurun): bash profiles/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2>> $out/${tag}_bench.err; echo "bench rc=$?"
python profiles/phases.py < $out/${tag}_bench.json
timeout 300 python bench.py --mode sample --steps 5 --warmup 2 > $out/${tag}_bench_sample.json 2>> $out/${tag}_bench.err
timeout 300 python bench.py --workload midi5shot_v4708_t256_h1024 --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_midi.json 2>> $out/${tag}_bench.err
timeout 300 python profiles/bench_softmax_grad.py 2,2,6 0,32,6 0,128,6 > $out/${tag}_softmax_grad.log 2>&1
# launch list of the bench command itself (cold-cache, serialised: shares, not absolutes).  FSMG_COOP=0: ncu cannot replay a
# cooperative launch that also has a cluster dimension (the pair-mode backward kernel)
FSMG_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_launches.log 2>&1
# --set full captures (one eager step): GEMM-core instantiations of the projection, the recurrent kernels, the softmax-grad pass
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 3 -f -o $out/${tag}_gemm \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_gemm.log 2>&1
FSMG_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o $out/${tag}_lstm \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_lstm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_grad -s 2 -c 1 -f -o $out/${tag}_softmax \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_softmax.log 2>&1
ls -la $out | grep ${tag}_
