#!/bin/bash
# usage: tools/ab_bench.sh workload libA libB [reps]  -- alternates two prebuilt libraries on one box
w=$1; a=$2; b=$3; reps=${4:-2}
for i in $(seq $reps); do
  for v in $a $b; do
    cp $v thunder_speech_b200/libthunder_b200.so
    echo -n "$v $w: "
    timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'top', r['kernel'], round(r['achieved']))"
  done
done
