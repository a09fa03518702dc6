set -x
mkdir -p gpurun_out
# synchronisation pass and DC paths: consumed bits by PRMT (config 5 and the multi-lane decode of small batches)
python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r6u_bench_config5.json 2> gpurun_out/r6u_bench.err; echo "bench5 rc=$?"
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6u_bench.json 2>> gpurun_out/r6u_bench.err; echo "bench rc=$?"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r6u_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6u_gputest.log
