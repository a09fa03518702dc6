set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29821 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r6n_reference_arm_2gpu.json 2> gpurun_out/r6n_reference_arm_2gpu.err; echo "ref rc=$?"
$TR --master-port 29822 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r6n_bench_2gpu.json 2> gpurun_out/r6n_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r6n_bench_2gpu.err
$TR --master-port 29823 bench.py --gpus 2 --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6n_bench_2gpu_config5.json 2> gpurun_out/r6n_bench_2gpu_config5.err; echo "bench5 rc=$?"
python -m pytest tests -m gpu -x -q -k "two_device or second_device or devices" > gpurun_out/r6n_gputest.log 2>&1; tail -2 gpurun_out/r6n_gputest.log
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r6n_*.json')):
    try:
        d=json.load(open(f)); print(f, d.get('impl'), round(d['value']), round(d.get('ms_per_step',0),2), 'n_gpus', d.get('n_gpus'), 'e2e', round(d['e2e']['value']), 'table', d.get('packet_table_ok'))
    except Exception as e: print(f, 'ERR', e)
"
