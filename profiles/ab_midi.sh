#!/bin/bash
# cfg3 (MIDI, H=1024, one episode) under recurrent-kernel variants
tag=$1; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1)); echo "=== midi variant $i: $v"
  env $v timeout 300 python bench.py --workload midi5shot_v4708_t256_h1024 --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_midi_$i.json 2>$out/${tag}_midi_$i.err; python profiles/phases.py < $out/${tag}_midi_$i.json
done
