set -x
mkdir -p gpurun_out
# lean VLC kernels: warps per CTA (7 / 8 / 11) chosen per launch for the fewest warps on the busiest SM
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r6z_bench.json 2> gpurun_out/r6z_bench.err; echo "bench rc=$?"
python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r6z_bench_config5.json 2>> gpurun_out/r6z_bench.err; echo "bench5 rc=$?"
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6z_bench_config4.json 2>> gpurun_out/r6z_bench.err; echo "bench4 rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r6z_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r6z_gputest.log
