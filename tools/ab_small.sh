#!/bin/bash
# A/B of the small-footprint (co-resident) kernel variants: option small = 0 (off) / 2 (always), per batch size.
cd "$(dirname "$0")/.."
for m in 0 2; do
  echo "== small=$m"
  THUNDER_B200_OPTIONS="small=$m" CHAINS=1 python tools/ab_chains.py quartznet15x5 ${QN_BATCHES:-16 32 64 128 256} 2>&1 | grep chains
  THUNDER_B200_OPTIONS="small=$m" CHAINS=1 python tools/ab_chains.py citrinet1024 ${CN_BATCHES:-16 32 128} 2>&1 | grep chains
done
