set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r5a_gputest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5a_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r5a_bench_new.json 2> gpurun_out/r5a_bench_new.err; echo "bench new rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt encode_rounds=1 --opt decode_tokens16=0 > gpurun_out/r5a_bench_old.json 2> gpurun_out/r5a_bench_old.err; echo "bench old rc=$?"
python profiles/prof_target.py 65536 0 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2|k_vlc_tokens16|k_idct16" -s 3 -c 3 -o gpurun_out/r5a_prof -f python profiles/prof_target.py 65536 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r5a_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python -c "
import json
for f in ('new','old'):
    try:
        d=json.load(open('gpurun_out/r5a_bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['kernels_ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"
