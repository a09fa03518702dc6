set -x
mkdir -p gpurun_out
# final kernels of round 2 (encoder as r6j; k_idct16 scatter without mask; token pass: parked lanes out of the long-code path, PRMT, bit-select):
# k_idct16: divisions by multiply, unconditional stores): ncu --set full of the config-2 kernels at 100 000 frames (second
# pass of prof_target) and the launch list of bench.py -- each after the same command has exited 0 without ncu
python profiles/prof_target.py 100000 0 > gpurun_out/plain_r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2|k_compact|k_unstuff|k_vlc_tokens_lean|k_idct16" -s 5 -c 8 -o gpurun_out/r6r_prof -f python profiles/prof_target.py 100000 0 > gpurun_out/ncu_full_r.log 2>&1; echo "ncu full rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r6r_bench_plain.json 2> gpurun_out/plain2_r.log && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r6r_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/ncu_launch_r.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r6r
