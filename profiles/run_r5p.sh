set -x
mkdir -p gpurun_out
# final state of round 2: ncu --set full of the config-2 kernels at 100 000 frames (second pass of prof_target), of the
# synchronisation pass (16 384 frames on 8 lanes each), and the launch list of bench.py -- each after the same command
# has exited 0 without ncu
python profiles/prof_target.py 100000 0 > gpurun_out/plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2|k_compact|k_unstuff|k_vlc_tokens_lean|k_idct16" -s 5 -c 5 -o gpurun_out/r5p_prof -f python profiles/prof_target.py 100000 0 > gpurun_out/ncu_full_p.log 2>&1; echo "ncu full rc=$?"
python profiles/prof_target.py 16384 3 > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_vlc_sync" -s 1 -c 1 -o gpurun_out/r5p_sync -f python profiles/prof_target.py 16384 3 > gpurun_out/ncu_full_s.log 2>&1; echo "ncu sync rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r5p_bench_plain.json 2> gpurun_out/plain2.log && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r5p_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r5p
