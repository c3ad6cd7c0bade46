# A/B of the persistent Toeplitz kernel inside the full model (graph replay): bash tools/ab_dw_persist.sh
for rep in 1 2; do
for opt in 0 2; do
  for b in 32 256; do
    echo -n "dw_persist=$opt quartznet15x5 batch=$b: "; THUNDER_B200_OPTIONS=dw_persist=$opt python bench.py --no-also --no-cpu-baseline --batch $b --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items()}, d['clocks']['sm_mhz'])"
  done
done
done
for opt in 0 2; do
  echo -n "dw_persist=$opt citrinet batch=16: "; THUNDER_B200_OPTIONS=dw_persist=$opt python bench.py --workload citrinet1024 --no-cpu-baseline --batch 16 --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items()})"
done
