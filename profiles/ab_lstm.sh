#!/bin/bash
# A/B of recurrent-kernel schedules: in-kernel timelines (FSMG_TRACE) + bench phases per variant.  usage: bash profiles/ab_lstm.sh <tag> "<env A>" "<env B>" ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1))
  echo "=== variant $i: $v" | tee -a $out/${tag}_ab.log
  env $v FSMG_TRACE=1 timeout 300 python profiles/profile_step.py 1 > $out/${tag}_trace_$i.log 2>&1; grep -A3 "fsmg trace" $out/${tag}_trace_$i.log | grep -v "^--" | tee -a $out/${tag}_ab.log; tail -1 $out/${tag}_trace_$i.log | tee -a $out/${tag}_ab.log
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_ab_$i.json 2> $out/${tag}_ab_$i.err; echo "rc=$?" | tee -a $out/${tag}_ab.log
  python profiles/phases.py < $out/${tag}_ab_$i.json | tee -a $out/${tag}_ab.log
done
