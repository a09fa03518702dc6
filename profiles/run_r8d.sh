set -x
mkdir -p gpurun_out
# k_idct16: coefficient words of a thread as one 144-byte row (8 x STS.128 to clear, 8 x LDS.128 to read) instead of a 32-word column
python bench.py --no-cpu-baseline > gpurun_out/r8d_bench.json 2> gpurun_out/r8d_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r8d_bench.err
python -m pytest tests -m gpu -x -q -k "decode or sp5x or mjpeg or config or golden or dropin or fixture" > gpurun_out/r8d_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r8d_gputest.log
python -c "
import json
d=json.load(open('gpurun_out/r8d_bench.json')); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
