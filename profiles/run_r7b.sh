set -x
mkdir -p gpurun_out
# the committed kernels at the end of round 2 (r7a state): ncu --set full at 100 000 frames and the launch list of bench.py
# k_idct16: divisions by multiply, unconditional stores): ncu --set full of the config-2 kernels at 100 000 frames (second
# pass of prof_target) and the launch list of bench.py -- each after the same command has exited 0 without ncu
python profiles/prof_target.py 100000 0 > gpurun_out/plain_b7.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2|k_compact|k_unstuff|k_vlc_tokens_lean|k_idct16" -s 5 -c 8 -o gpurun_out/r7b_prof -f python profiles/prof_target.py 100000 0 > gpurun_out/ncu_full_b7.log 2>&1; echo "ncu full rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r7b_bench_plain.json 2> gpurun_out/plain2_b7.log && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r7b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/ncu_launch_b7.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r7b
