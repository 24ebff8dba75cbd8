mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_gpu_gather.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/m2_tests.log 2>&1; echo GATHER_TEST_EXIT $?; tail -3 gpurun_out/m2_tests.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; echo BENCH2_EXIT $?; tail -2 gpurun_out/m2_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/m2_bench.json'))
print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']), 'e2e', d['e2e']['value'], 'gather', d.get('gather'), 'repro', d.get('bitwise_reproducible_across_steps'))
PY
