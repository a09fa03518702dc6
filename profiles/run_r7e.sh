set -x
mkdir -p gpurun_out
# the committed binary under torchrun at 2 GPUs, launched as the driver launches it
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29831 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r7e_bench_2gpu.json 2> gpurun_out/r7e_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r7e_bench_2gpu.err
$TR --master-port 29832 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r7e_reference_arm_2gpu.json 2> gpurun_out/r7e_reference_arm_2gpu.err; echo "ref rc=$?"
