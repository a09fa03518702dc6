set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "adpcm" > gpurun_out/r5g_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5g_gputest.log
for f in 1 2; do
  python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --opt adpcm_form=$f > gpurun_out/r5g_bench_config3_form$f.json 2> gpurun_out/r5g_bench_config3_form$f.err; echo "bench adpcm form $f rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5g_bench*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()}, d['roofline']['frac'], d.get('audit'))
    except Exception as e: print(f, 'ERR', e)
"
