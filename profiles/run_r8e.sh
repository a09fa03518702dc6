set -x
mkdir -p gpurun_out
# k_idct16 with the multiply-add chain transform at 6 / 7 / 8 CTAs per SM (80 / 72 / 64 registers, 0 / 28 / 92 bytes of spills)
# (variants/libamvcuda_*.so: builds of the named source states made for this A/B with "make" and copied aside; not kept in the tree)
cp amv-codec-tools_b200/lib/libamvcuda.so /tmp/lib7.so
for n in 6 8 7; do
  if [ $n = 7 ]; then cp /tmp/lib7.so amv-codec-tools_b200/lib/libamvcuda.so; else cp variants/libamvcuda_idct$n.so amv-codec-tools_b200/lib/libamvcuda.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8e_bench_idct$n.json 2> gpurun_out/r8e_bench.err; echo "bench $n rc=$?"
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r8e_*.json')):
    d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
