set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "two_devices" > gpurun_out/r5e_twodev.log 2>&1; echo "two-device test rc=$?"; tail -3 gpurun_out/r5e_twodev.log
for c in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 5 --warmup 3 > gpurun_out/r5e_bench_2gpu_config$c.json 2> gpurun_out/r5e_bench_2gpu_config$c.err; echo "bench 2gpu config $c rc=$?"; tail -2 gpurun_out/r5e_bench_2gpu_config$c.err
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5e_bench*.json')):
    try:
        d=json.load(open(f)); print(f, d['metric'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'table', d.get('packet_table_ok'), d.get('audit'))
    except Exception as e: print(f, 'ERR', e)
"
