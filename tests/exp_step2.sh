run() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --profile-out gpurun_out/prof_$tag.json 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value']), round(d['ms_per_step'],3))"; }
run base SVK_GATHER3_MMA=1
run x1 SVK_GATHER3_MMA=2 SVK_EPI2_MODE=7 SVK_EPI2_GROUPS=3 SVK_GATHER3_STAGES=12
run x3 SVK_GATHER3_MMA=2 SVK_EPI2_MODE=1 SVK_EPI2_GROUPS=3 SVK_GATHER3_GROUPS=3 SVK_GATHER3_STAGES=12
run x4 SVK_GATHER3_MMA=2 SVK_GATHER3_GROUPS=3 SVK_GATHER3_STAGES=12
run x5 SVK_GATHER3_MMA=2 SVK_GATHER3_STAGES=12
