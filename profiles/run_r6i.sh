set -x
mkdir -p gpurun_out
# encoder: AC loop one coefficient ahead (two register sets)
for i in 1 2; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6i_bench_$i.json 2> gpurun_out/r6i_bench.err; echo "bench rc=$?"
done
