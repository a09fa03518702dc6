set -x
mkdir -p gpurun_out
# k_idct16 scatter: validity of a group's eight tokens as one bit mask (B) against an add and a compare per token (A)
# (variants/libamvcuda_*.so: builds of the named source states made for this A/B with "make" and copied aside; not kept in the tree)
cp amv-codec-tools_b200/lib/libamvcuda.so /tmp/libB.so
for v in A B A B; do
  cp variants/libamvcuda_$v.so amv-codec-tools_b200/lib/libamvcuda.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8g_bench_$v.json 2> gpurun_out/r8g_bench.err; echo "bench $v rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r8g_bench_$v.json')); print('$v', round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
done
cp /tmp/libB.so amv-codec-tools_b200/lib/libamvcuda.so
python -m pytest tests -m gpu -x -q -k "decode or sp5x or config or golden" > gpurun_out/r8g_gputest.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r8g_gputest.log
