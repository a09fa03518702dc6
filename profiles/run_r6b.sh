set -x
mkdir -p gpurun_out
# encoder after the Huffman-loop cuts (shifted-out masks, size from leading zeros, sign-test bit writer, macroblock row by multiply)
for f in 4 2 8; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6b_bench_form$f.json 2> gpurun_out/r6b_bench_form$f.err; echo "form $f rc=$?"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encode" > gpurun_out/r6b_gputest_encode.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6b_gputest_encode.log
