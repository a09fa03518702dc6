set -x
mkdir -p gpurun_out
# simple_idct passes as multiply-add chains (36 instead of 44 instructions per 8-point pass); token workspace of the 16-bit pass halved
python bench.py --no-cpu-baseline > gpurun_out/r8c_bench.json 2> gpurun_out/r8c_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r8c_bench.err
python -m pytest tests -m gpu -x -q -k "decode or sp5x or mjpeg or config or golden or dropin or fixture" > gpurun_out/r8c_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r8c_gputest.log
python -c "
import json
d=json.load(open('gpurun_out/r8c_bench.json')); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
