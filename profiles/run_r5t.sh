set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5t_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5t_gputest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5t_bench.json 2> gpurun_out/r5t_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r5t_bench.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5t_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'] if 'roofline' in d else '')
    except Exception as e: print(f, 'ERR', e)
"
python __graft_entry__.py smoke 2>&1 | tail -1
