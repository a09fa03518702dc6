set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5m_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5m_gputest.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "dropin_matches_reference_codecs" 2>&1 | grep "drop-in frames" > gpurun_out/r5m_dropin_fps.txt; cat gpurun_out/r5m_dropin_fps.txt
for t in 0 2; do
  python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline --opt decode_token_pass=$t > gpurun_out/r5m_bench_config5_pass$t.json 2> gpurun_out/r5m_bench_config5_pass$t.err; echo "bench5 pass $t rc=$?"
done
python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5m_bench_config4.json 2> gpurun_out/r5m_bench_config4.err; echo "bench4 rc=$?"
python bench.py > gpurun_out/r5m_bench.json 2> gpurun_out/r5m_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r5m_bench.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5m_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['roofline']['frac'] if 'roofline' in d else '')
    except Exception as e: print(f, 'ERR', e)
"
python __graft_entry__.py smoke 2>&1 | tail -1
