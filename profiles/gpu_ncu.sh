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
tail -2 $out/${tag}_ncu_*.log
ls -la $out | grep ${tag}_
