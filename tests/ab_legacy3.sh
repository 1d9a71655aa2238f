# A/B: forward kernels of the 32/64-channel stages with two issuing warps / three groups / 12 stages (default) vs the round-1 form
for i in 1 2; do
  for cfg in "SVK_GATHER3_MMA=1 SVK_GATHER3_GROUPS=2 SVK_GATHER3_STAGES=6" "SVK_X=0"; do
    env $cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
  done
done
