set -x
mkdir -p gpurun_out
# k_unstuff: RSTn flags only when an FF is followed by something other than 00 / FF; the byte in front of a unit by shuffle instead of a global byte load
python bench.py --no-cpu-baseline > gpurun_out/r8b_bench.json 2> gpurun_out/r8b_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r8b_bench.err
python -m pytest tests -m gpu -x -q -k "decode or unstuff or corrupt or sp5x or mjpeg or config" > gpurun_out/r8b_gputest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r8b_gputest.log
python -c "
import json
d=json.load(open('gpurun_out/r8b_bench.json')); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
