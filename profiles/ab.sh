#!/bin/bash
# A/B helper for one GPU-box visit: runs bench.py once per environment variant and prints ms/step + phases.
#   bash profiles/ab.sh <tag> "VAR=1 VAR2=0" "VAR=2" ...      ("-" = defaults)
tag=$1; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1))
  [ "$v" = "-" ] && v=""
  echo "=== variant $i: ${v:-defaults}" | tee -a $out/${tag}_ab.log
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${AB_EXTRA:---no-extra-configs} > $out/${tag}_ab_$i.json 2> $out/${tag}_ab_$i.err
  rc=$?
  if [ $rc -ne 0 ]; then echo "rc=$rc"; tail -5 $out/${tag}_ab_$i.err; fi | tee -a $out/${tag}_ab.log
  python profiles/phases.py < $out/${tag}_ab_$i.json 2>&1 | tee -a $out/${tag}_ab.log
done
