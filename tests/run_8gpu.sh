TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2> gpurun_out/r02_8gpu_cfg3.err | tail -1 > gpurun_out/r02_bench_cfg3_8gpu.json
$TR --master-port 29802 bench.py --gpus 8 --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_8gpu_cfg4.err | tail -1 > gpurun_out/r02_bench_cfg4_8gpu.json
$TR --master-port 29803 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 1 2> gpurun_out/r02_8gpu_cfg5.err | tail -1 > gpurun_out/r02_bench_cfg5_8gpu.json
for f in cfg3 cfg4 cfg5; do python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${f}_8gpu.json')); print('$f', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d.get('checks'), d['e2e']['value'])" || tail -5 gpurun_out/r02_8gpu_$f.err; done
