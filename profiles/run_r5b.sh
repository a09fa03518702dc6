set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5b_gputest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5b_gputest.log
for v in "decode_token_pass=1" "decode_token_pass=2" "decode_token_pass=0" "encode_rounds=3"; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-frames 4096 --audit 0 --opt $v > gpurun_out/r5b_bench_$v.json 2> gpurun_out/r5b_bench_$v.err; echo "bench $v rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5b_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
"
