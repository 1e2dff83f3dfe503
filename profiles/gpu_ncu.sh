#!/bin/bash
# ncu evidence for one round: launch list of the bench command + `--set full` captures of the hot kernels (cfg 2 GEMMs / recurrent /
# softmax-grad, cfg 3 recurrent, cfg 5 decode step).  FSMG_COOP=0: ncu cannot replay a cooperative launch that carries a cluster dimension.
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
FSMG_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_launches.log 2>&1
FSMG_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o $out/${tag}_lstm \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_lstm.log 2>&1
FSMG_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o $out/${tag}_lstm_midi \
    python profiles/profile_step.py 1 midi5shot_v4708_t256_h1024 > $out/${tag}_ncu_lstm_midi.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 3 -f -o $out/${tag}_gemm \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_grad -s 2 -c 1 -f -o $out/${tag}_softmax \
    python profiles/profile_step.py 1 > $out/${tag}_ncu_softmax.log 2>&1
FSMG_SAMPLE_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_kernel|sample_cell|argmax_rows" -s 9 -c 4 -f -o $out/${tag}_sample \
    python profiles/profile_sample.py 6 > $out/${tag}_ncu_sample.log 2>&1
# the reports themselves exceed gpurun's 64 MiB return limit: summarise on the box, keep the raw-page CSV (one row per kernel), drop the reports
for r in lstm lstm_midi gemm softmax sample; do
  ncu -i $out/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_$r.raw.csv 2>/dev/null
done
python profiles/summarize_ncu.py $out/${tag}_gemm.ncu-rep $out/${tag}_lstm.ncu-rep $out/${tag}_softmax.ncu-rep $out/${tag}_lstm_midi.ncu-rep $out/${tag}_sample.ncu-rep > $out/${tag}_ncu_summary.md 2>&1
python profiles/summarize_launches.py $out/${tag}_launches.csv > $out/${tag}_launches_summary.md 2>&1
rm -f $out/${tag}_*.ncu-rep
for f in $out/${tag}_ncu_*.log; do tail -n 2 $f; done
ls -la $out | grep ${tag}_
