# usage: bash profiles/run_scale.sh N   (on a box with N GPUs): PCIe probe at 1..N ranks, then bench.py at N ranks for every config
set -x
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc; grep -m1 "model name" /proc/cpuinfo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in ${PROBE_N:-1 2 4 8}; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port $((29600+n)) profiles/pcie_probe_multi.py 512 2>/dev/null | tail -1 > gpurun_out/${TAG:-r5h}_pcie_${n}gpu.json; cat gpurun_out/${TAG:-r5h}_pcie_${n}gpu.json
done
for c in ${CONFIGS:-2 4 5 3}; do
  $TR --nproc-per-node $N --master-port $((29700+c)) bench.py --gpus $N --config $c --steps 8 --warmup 3 > gpurun_out/${TAG:-r5h}_bench_${N}gpu_config$c.json 2> gpurun_out/${TAG:-r5h}_bench_${N}gpu_config$c.err; echo "bench ${N}gpu config $c rc=$?"; tail -1 gpurun_out/${TAG:-r5h}_bench_${N}gpu_config$c.err
done
# the e2e leg with DMA'd instead of zero-copy packets
$TR --nproc-per-node $N --master-port 29790 bench.py --gpus $N --config 2 --steps 8 --warmup 3 --no-cpu-baseline --opt host_zero_copy_packets=0 > gpurun_out/${TAG:-r5h}_bench_${N}gpu_config2_dma_packets.json 2> gpurun_out/${TAG:-r5h}_bench_${N}gpu_config2_dma_packets.err; echo "bench dma packets rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG:-r5h}_bench_${N}gpu*.json')):
    try:
        d=json.load(open(f)); print(f, d['metric'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'table', d.get('packet_table_ok'), 'audit', d.get('audit',{}).get('ok'), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
    except Exception as e: print(f, 'ERR', e)
"
