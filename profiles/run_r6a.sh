set -x
mkdir -p gpurun_out
# A/B of the encoder's transform forms (encode_rounds 2 = factorised, 4..8 = regrouped for the two integer pipes)
for f in 2 4 5 6 7 8 3; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt encode_rounds=$f > gpurun_out/r6a_bench_form$f.json 2> gpurun_out/r6a_bench_form$f.err; echo "form $f rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r6a_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d.get('kernels_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']), d['audit'] if 'audit' in d else '')
    except Exception as e: print(f, 'ERR', e)
"
