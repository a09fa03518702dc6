set -x
mkdir -p gpurun_out
# encoder: stuffing loop without the per-byte "inside the segment?" tests (the short last word's stray bytes land behind the payload, EOI written after the loop)
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6w_bench.json 2> gpurun_out/r6w_bench.err; echo "bench rc=$?"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "encode or config or shape" > gpurun_out/r6w_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6w_gputest.log
