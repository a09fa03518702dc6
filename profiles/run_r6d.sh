set -x
mkdir -p gpurun_out
# encoder: pixel rows requested one round ahead (forms 9 / 10) against forms 4 / 8
for f in 4 8 9 10; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6d_bench_form$f.json 2> gpurun_out/r6d_bench_form$f.err; echo "form $f rc=$?"
done
