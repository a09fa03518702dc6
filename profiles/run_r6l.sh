set -x
mkdir -p gpurun_out
# k_idct16 as a block-stride loop: resident grid of 6 / 12 CTAs per SM against one CTA per 128 blocks (0); quantiser mask dropped
for c in 0 6 12; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt idct_ctas_per_sm=$c > gpurun_out/r6l_bench_ctas$c.json 2> gpurun_out/r6l_bench.err; echo "ctas $c rc=$?"
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r6l_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6l_gputest.log
