set -x
mkdir -p gpurun_out
# encoder: seven CTAs of three warps per SM (form 11) against form 4
for f in 11 4; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6f_bench_form${f}.json 2> gpurun_out/r6f_bench.err; echo "form $f rc=$?"
done
