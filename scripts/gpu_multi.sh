# multi-GPU bench: torchrun, one rank per GPU (the driver's launch line)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L
timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_multi.log 2>&1; echo TEST_EXIT $?; tail -2 gpurun_out/test_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo BENCH_EXIT $?
cat gpurun_out/bench_${N}gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_${N}gpu_ref.json 2> gpurun_out/bench_${N}gpu_ref.err; echo REF_EXIT $?
cat gpurun_out/bench_${N}gpu_ref.json | cut -c1-600
