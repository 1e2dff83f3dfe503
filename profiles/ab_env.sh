#!/bin/bash
# generic A/B of environment variants on the bench line (10 steps).  usage: bash profiles/ab_env.sh <tag> "<env A>" "<env B>" ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1)); echo "=== variant $i: $v"
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_env_$i.json 2> $out/${tag}_env_$i.err; echo "rc=$?"
  python profiles/phases.py < $out/${tag}_env_$i.json
done
