set -x
mkdir -p gpurun_out
# k_unstuff: the flush and the clearing of the stage written out as two predicated 128-bit moves per thread instead of the general loops
timeout 25 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 2048 --audit 0 > gpurun_out/r8m_bench.json 2> gpurun_out/r8m_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r8m_bench.json')); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d.get('kernels_ms_per_step',{}).items()})
"
timeout 32 python -m pytest tests -m gpu -x -q -k "decode or corrupt or sp5x" > gpurun_out/r8m_gputest.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r8m_gputest.log
