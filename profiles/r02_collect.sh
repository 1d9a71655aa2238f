#!/bin/bash
# r02 evidence: run on a B200 under gpurun (one GPU).  Every ncu pass runs only after the same command exited 0 without ncu.
#   bash profiles/r02_collect.sh
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
# 1. sustained number (>= 4 s timed region) + clocks
python bench.py --steps 500 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_sustained.json 2> gpurun_out/r02_bench_sustained.err
echo "sustained rc=$?"
# 2. kernel timeline of one step (CUPTI), after PDL
python tests/prof_timeline.py --out gpurun_out/r02_timeline_after_pdl.json > gpurun_out/r02_timeline_after_pdl.log 2>&1
echo "timeline rc=$?"
# 3. plain run, then the ncu passes of the SAME command
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv \
    --log-file gpurun_out/r02_ncu_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
full() {   # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/r02_full_$1 $B > gpurun_out/r02_full_$1.log 2>&1
  echo "full $1 rc=$?"
}
full gather3 conv_tc_gather3_kernel 78 26
full gather2 conv_tc_gather2_kernel 123 41
full wgrad9 conv_tc_wgrad9_kernel 39 13
full wgradr conv_tc_wgradr_kernel 48 16
full bn_bwd_apply bn_bwd_apply_kernel 99 33
full bn_train_act bn_train_act_kernel 99 33
full stem "stem_(fwd|wgrad)_kernel" 6 2
full aam "aam_ce_" 12 4
ls -la gpurun_out/*.ncu-rep
