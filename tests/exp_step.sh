# training-step A/B of the resident-filter conv switches (same box, alternating)
run() { echo "== $*"; env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); kb=d['kernel_breakdown_ms']
print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in kb.items() if 'conv2d_fwd' in k or 'dgrad' in k})"; }
for i in 1 2; do
run SVK_GATHER3_MMA=1
run SVK_GATHER3_MMA=2 SVK_EPI2_MODE=7 SVK_EPI2_GROUPS=3 SVK_GATHER3_STAGES=12
run SVK_GATHER3_MMA=2 SVK_EPI2_MODE=7 SVK_EPI2_GROUPS=3 SVK_GATHER3_STAGES=12 SVK_SINGLE_HALO_FWD32=1
run SVK_GATHER3_MMA=2 SVK_EPI2_MODE=1 SVK_EPI2_GROUPS=3 SVK_GATHER3_GROUPS=3 SVK_GATHER3_STAGES=12
run SVK_GATHER3_MMA=2 SVK_EPI2_MODE=5 SVK_EPI2_GROUPS=3 SVK_GATHER3_STAGES=12
done
