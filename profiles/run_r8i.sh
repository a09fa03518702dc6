set -x
mkdir -p gpurun_out
# final state of round 2: whole GPU suite, smoke, default bench, reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/r8i_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r8i_gputest.log
python __graft_entry__.py smoke > gpurun_out/r8i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r8i_smoke.log
python bench.py > gpurun_out/r8i_bench.json 2> gpurun_out/r8i_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r8i_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r8i_reference_arm.json 2> gpurun_out/r8i_reference_arm.err; echo "ref rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r8i_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'] if 'roofline' in d else '', 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']))
    except Exception as e: print(f, 'ERR', e)
"
