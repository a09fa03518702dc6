set -x
mkdir -p gpurun_out
# ADPCM decode: 32-byte tiles (7 CTAs of 4 warps per SM instead of 4)
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r7f_bench_config3.json 2> gpurun_out/r7f_bench.err; echo "bench3 rc=$?"
python -m pytest tests -m gpu -x -q -k "adpcm" > gpurun_out/r7f_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r7f_gputest.log
