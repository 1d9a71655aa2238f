# usage: bash tests/ab_lib.sh <variant-lib-name> [rounds]  — alternates bench runs of svk/libsvk_<name>.so and svk/libsvk.so
V=$1; N=${2:-2}
for i in $(seq 1 $N); do
  SVK_LIB_PATH=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk_$V.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:v for k,v in d['kernel_breakdown_ms'].items() if 'bn_' in k})"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:v for k,v in d['kernel_breakdown_ms'].items() if 'bn_' in k})"
done
