# round 2, call 6 (2 GPUs): one process over both devices (tests), torchrun bench at N=2, configs 4 and 5 at reduced size
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; nvidia-smi topo -m 2>/dev/null | head -8
timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_gather.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/f_tests.log 2>&1; echo TEST_EXIT $?; grep -h "MULTI_OK\|passed\|failed\|skipped" gpurun_out/f_tests.log | tail -5; tail -30 gpurun_out/f_tests.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err; echo "BENCH2 exit $?"; tail -5 gpurun_out/f_bench2.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/f_bench2.json'))
    print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']))
    print('e2e', json.dumps(d['e2e'])[:800])
    print('e2e_s16', json.dumps(d['e2e_s16'])[:300])
    print('gather', d['gather'], 'per_rank', d['per_rank'])
except Exception as e:
    print('bench line unreadable', e)
PY
BLISS_CFG5_SONGS=3000 timeout 900 python bench.py --config 5 --gpus 2 > gpurun_out/f_cfg5.json 2> gpurun_out/f_cfg5.err; echo "CFG5 exit $?"; tail -5 gpurun_out/f_cfg5.err | cut -c1-400; cut -c1-3000 gpurun_out/f_cfg5.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_config4.py --total-songs 6000 > gpurun_out/f_cfg4.json 2> gpurun_out/f_cfg4.err; echo "CFG4 exit $?"; tail -5 gpurun_out/f_cfg4.err | cut -c1-400; cut -c1-1500 gpurun_out/f_cfg4.json
