for i in 1 2; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
  python bench.py --cuda-graph --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"
done
