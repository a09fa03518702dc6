set -x
mkdir -p gpurun_out
# k_idct16 at 9 / 10 CTAs per SM (56 / 48 registers), then the committed state (8 CTAs per SM): whole GPU suite, smoke, default bench, reference arm, configs 3 / 4 / 5
# (variants/libamvcuda_*.so: builds of the named source states made for this A/B with "make" and copied aside; not kept in the tree)
cp amv-codec-tools_b200/lib/libamvcuda.so /tmp/lib8.so
for n in 9 10; do
  cp variants/libamvcuda_idct$n.so amv-codec-tools_b200/lib/libamvcuda.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8f_bench_idct$n.json 2> gpurun_out/r8f_bench.err; echo "bench $n rc=$?"
done
cp /tmp/lib8.so amv-codec-tools_b200/lib/libamvcuda.so
python -m pytest tests -m gpu -x -q > gpurun_out/r8f_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r8f_gputest.log
python __graft_entry__.py smoke > gpurun_out/r8f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r8f_smoke.log
python bench.py > gpurun_out/r8f_bench.json 2> gpurun_out/r8f_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r8f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r8f_reference_arm.json 2> gpurun_out/r8f_reference_arm.err; echo "ref rc=$?"
for c in 3 4 5; do
  python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/r8f_bench_config$c.json 2> gpurun_out/r8f_bench_config$c.err; echo "bench$c rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r8f_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'] if 'roofline' in d else '', 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']))
    except Exception as e: print(f, 'ERR', e)
"
