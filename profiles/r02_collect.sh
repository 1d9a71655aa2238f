#!/bin/bash
# r02 evidence: run on a B200 under gpurun (one GPU).  Every ncu pass runs only after the same command exited 0 without ncu.
# Everything left in gpurun_out/ is small (CSV / JSON; two single-kernel .ncu-rep files): gpurun copies back <= 64 MiB.
#   bash profiles/r02_collect.sh
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_reads.sum,l1tex__data_bank_writes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,launch__registers_per_thread,launch__shared_mem_per_block_dynamic"
# 1. sustained number (>= 4 s timed region) + clocks
python bench.py --steps 500 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_sustained.json 2> gpurun_out/r02_bench_sustained.err
echo "sustained rc=$?"
# 2. kernel timeline of one step (CUPTI), after PDL
python tests/prof_timeline.py --out gpurun_out/r02_timeline_after_pdl.json > gpurun_out/r02_timeline_after_pdl.log 2>&1
echo "timeline rc=$?"
# 3. plain run, then the ncu passes of the SAME command
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "plain run failed"; exit 1; }
# every launch of two training steps with its time, tensor-pipe %, DRAM bytes / %, L2 %, L1 %, shared-memory wavefronts
ncu --metrics $M --clock-control none -s 700 -c 520 --csv --log-file gpurun_out/r02_ncu_step_metrics.csv $B > gpurun_out/r02_ncu_step.log 2>&1
echo "step metrics rc=$?"
full() {   # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02_full_$1 $B > gpurun_out/r02_full_$1.log 2>&1
  echo "full $1 rc=$?"
  ncu -i gpurun_out/r02_full_$1.ncu-rep --page raw --csv > gpurun_out/r02_full_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_full_$1.ncu-rep --page details --csv > gpurun_out/r02_full_$1.details.csv 2>/dev/null
}
full gather3_fwd32 conv_tc_gather3_kernel 78
full gather2_fwd128 conv_tc_gather2_kernel 123
ls -la gpurun_out/ | tail -20
du -sh gpurun_out
