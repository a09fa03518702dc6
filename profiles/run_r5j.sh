set -x
mkdir -p gpurun_out
for k in 0 2 4 8 16; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-frames 4096 --audit 0 --opt decode_overlap=$k > gpurun_out/r5j_bench_overlap$k.json 2> gpurun_out/r5j_bench_overlap$k.err; echo "bench overlap $k rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5j_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
"
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config2" 2>&1 | tail -2
