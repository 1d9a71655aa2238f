#!/bin/bash
# r02, final build: every launch of the training step under `ncu --metrics` (the same command exits 0 without ncu first).
#   bash profiles/r02_collect_final.sh     (one GPU, under gpurun; leaves CSV / JSON only)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_reads.sum,l1tex__data_bank_writes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,launch__registers_per_thread,launch__shared_mem_per_block_dynamic"
$B > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics $M --clock-control none -s 700 -c 330 --csv --log-file gpurun_out/r02f_ncu_step_metrics.csv $B > gpurun_out/r02f_ncu_step.log 2>&1
echo "step metrics rc=$?"
ls -la gpurun_out/r02f_* ; du -sh gpurun_out
