# fused peer gather: GPU tests of the exchange, the 2-process worker, then the N-GPU bench (auto = fused, NCCL comparator inside)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L; nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_gpu_gather.py -m gpu -q --tb=short -p no:cacheprovider -k two_processes > gpurun_out/test_gather.log 2>&1; echo TEST_EXIT $?; tail -4 gpurun_out/test_gather.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo BENCH_EXIT $?
cat gpurun_out/bench_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','gather','per_rank','gpu_launches')}); print(d['e2e'])"; tail -5 gpurun_out/bench_${N}gpu.err
