set -x
mkdir -p gpurun_out
# state after the re-entry cuts of round 2: the whole GPU suite, the default bench, the reference arm, the other configurations, smoke
python -m pytest tests -m gpu -x -q > gpurun_out/r7a_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r7a_gputest.log
python bench.py > gpurun_out/r7a_bench.json 2> gpurun_out/r7a_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r7a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r7a_reference_arm.json 2> gpurun_out/r7a_reference_arm.err; echo "ref rc=$?"
for c in 3 4 5; do
  python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/r7a_bench_config$c.json 2> gpurun_out/r7a_bench_config$c.err; echo "bench$c rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r7a_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'] if 'roofline' in d else '', 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']))
    except Exception as e: print(f, 'ERR', e)
"
python __graft_entry__.py smoke 2>&1 | tail -1
