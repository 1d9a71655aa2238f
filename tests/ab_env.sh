# usage: bash tests/ab_env.sh VAR=VALUE [rounds]   — alternates bench runs with and without the environment switch
KV=$1; N=${2:-2}
for i in $(seq 1 $N); do
  env $KV python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$KV', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
