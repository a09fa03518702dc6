set -x
mkdir -p gpurun_out
# encoder after pruning the rejected forms: default (form 4), 2, 8; encode parity tests
for f in 4 2 8; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6g_bench_form${f}.json 2> gpurun_out/r6g_bench.err; echo "form $f rc=$?"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encode" > gpurun_out/r6g_gputest_encode.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6g_gputest_encode.log
