set -x
mkdir -p gpurun_out
# lean token pass in CTAs of 11 warps (two per SM: at most 22 warps on an SM instead of 24 / 16)
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6x_bench.json 2> gpurun_out/r6x_bench.err; echo "bench rc=$?"
python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r6x_bench_config5.json 2>> gpurun_out/r6x_bench.err; echo "bench5 rc=$?"
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6x_bench_config4.json 2>> gpurun_out/r6x_bench.err; echo "bench4 rc=$?"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r6x_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6x_gputest.log
