set -x
mkdir -p gpurun_out
# encoder: bit writer / packer tails by one funnel shift, stuffing loop split into whole words (no per-byte tests) and the short last word
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6s_bench.json 2> gpurun_out/r6s_bench.err; echo "bench rc=$?"
python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r6s_bench_config5.json 2>> gpurun_out/r6s_bench.err; echo "bench5 rc=$?"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r6s_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6s_gputest.log
