#!/bin/bash
# A/B of library option sets on the captured forward: tools/ab_opts.sh "optset1" "optset2" ...   (QN_BATCHES / CN_BATCHES)
cd "$(dirname "$0")/.."
for o in "$@"; do
  echo "== $o"
  THUNDER_B200_OPTIONS="$o" CHAINS=1 python tools/ab_chains.py quartznet15x5 ${QN_BATCHES:-256} 2>&1 | grep chains
  if [ -n "$CN_BATCHES" ]; then THUNDER_B200_OPTIONS="$o" CHAINS=1 python tools/ab_chains.py citrinet1024 $CN_BATCHES 2>&1 | grep chains; fi
done
