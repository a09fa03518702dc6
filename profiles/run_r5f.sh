set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5f_gputest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5f_gputest.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "dropin_matches_reference_codecs" 2>&1 | grep "drop-in frames" > gpurun_out/r5f_dropin_fps.txt; cat gpurun_out/r5f_dropin_fps.txt
for f in 0 1 2; do
  python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --opt adpcm_form=$f > gpurun_out/r5f_bench_config3_form$f.json 2> gpurun_out/r5f_bench_config3_form$f.err; echo "bench adpcm form $f rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r5f_bench*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']), d.get('audit'))
    except Exception as e: print(f, 'ERR', e)
"
python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4096 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_adpcm" -s 6 -c 2 -o gpurun_out/r5f_adpcm -f python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4096 > gpurun_out/ncu_adpcm.log 2>&1; echo "ncu adpcm rc=$?"
