#!/bin/bash
# Round-2, session-2 GPU visit ("r3" files): whole GPU suite, the bench line the driver will run, ncu evidence for the
# fused-softmax-gradient kernels (launch list, --set full of the three projection GEMMs, per-instruction stall samples of the
# logits + LSE kernel), then A/B diagnostics.
# usage (under gpurun): bash profiles/gpu_visit_r3.sh <tag> [ab variants...]
tag=${1:-r3a}; shift
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python profiles/phases.py < $out/${tag}_bench.json
# launch list (cold-cache, serialised: shares only)
FSMG_COOP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_launches.log 2>&1
python profiles/summarize_launches.py $out/${tag}_launches.csv > $out/${tag}_launches_summary.md 2>&1
# the chunk's three GEMMs: logits + LSE (exp store), dH (XF = 1), dWs (XF = 2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 3 -f -o $out/${tag}_gemm \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_gemm.log 2>&1
ncu -i $out/${tag}_gemm.ncu-rep --page raw --csv > $out/${tag}_gemm.raw.csv 2>/dev/null
ncu -i $out/${tag}_gemm.ncu-rep --page source --csv --print-source sass > $out/${tag}_gemm.source.csv 2>/dev/null
python profiles/stall_summary.py $out/${tag}_gemm.source.csv 14 > $out/${tag}_stalls_summary.md 2>&1
[ "$KEEP_SOURCE_CSV" = "1" ] || rm -f $out/${tag}_gemm.source.csv
python profiles/summarize_ncu.py $out/${tag}_gemm.ncu-rep > $out/${tag}_ncu_summary.md 2>&1
rm -f $out/${tag}_*.ncu-rep
tail -n 3 $out/${tag}_ncu_gemm.log
[ $# -gt 0 ] && bash profiles/ab.sh $tag "$@"
ls -la $out | grep ${tag}_
