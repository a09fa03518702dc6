set -x
mkdir -p gpurun_out
# final kernels of round 2: ncu --set full at 100 000 frames (second pass of prof_target), then the launch list of bench.py -- each after the same command has exited 0 without ncu
timeout 60 python profiles/prof_target.py 100000 0 > gpurun_out/r8j_plain.log 2>&1 && timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2|k_compact|k_unstuff|k_vlc_tokens_lean|k_idct16" -s 5 -c 8 -o gpurun_out/r8j_prof -f python profiles/prof_target.py 100000 0 > gpurun_out/r8j_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 40 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8j_bench_plain.json 2> gpurun_out/r8j_plain2.log && timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r8j_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8j_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r8j
