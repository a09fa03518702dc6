set -x
mkdir -p gpurun_out
# the final binary under torchrun at 2 GPUs, launched as the driver launches it
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 60 $TR --master-port 29841 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r8l_bench_2gpu.json 2> gpurun_out/r8l_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r8l_bench_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r8l_bench_2gpu.json')); print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d.get('packet_table_ok'), d['n_gpus'])
"
