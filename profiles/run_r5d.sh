set -x
mkdir -p gpurun_out
echo skip tests
python bench.py > gpurun_out/r5d_bench.json 2> gpurun_out/r5d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r5d_bench.err
for c in 5 4; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r5d_bench_config$c.json 2> gpurun_out/r5d_bench_config$c.err; echo "bench config $c rc=$?"; tail -2 gpurun_out/r5d_bench_config$c.err
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5d_bench*.json')):
    try:
        d=json.load(open(f)); print(f, d['metric'], round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d.get('audit'))
    except Exception as e: print(f, 'ERR', e)
"
