for i in 1 2 3; do
  SVK_DISABLE_FUSED_HEAD=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unfused', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused  ', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
done
