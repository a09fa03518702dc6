set -x
mkdir -p gpurun_out
# encoder: next round's rows asked into the L2 at the start of the Huffman stage (form 10) against form 4
for f in 4 10 4 10; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6e_bench_form${f}_$RANDOM.json 2> gpurun_out/r6e_bench.err; echo "form $f rc=$?"
done
