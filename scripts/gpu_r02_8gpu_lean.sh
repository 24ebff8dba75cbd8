# round 2 (8 GPUs, one box), closing refresh on the final build: the bench line at N=8, BASELINE configs[3], the
# multi-device tests.  (gpu_r02_8gpu.sh is the full run: N=4, configs[4], the reference arm.)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/k_bench8.json 2> gpurun_out/k_bench8.err; echo "BENCH8 exit $?"; tail -2 gpurun_out/k_bench8.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/k_bench8.json'))
    print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']))
    print('e2e %.0f frac %.3f s16 %.0f cd %s' % (d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], d['e2e_s16']['value'], json.dumps(d.get('e2e_cd'))[:600]))
    print('gather', d['gather'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 600 $TR --master-port 29522 bench.py --config 4 --gpus 8 > gpurun_out/k_cfg4.json 2> gpurun_out/k_cfg4.err; echo "CFG4 exit $?"; cut -c1-900 gpurun_out/k_cfg4.json
timeout 600 python -m pytest tests/test_gpu_multidevice.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/k_tests.log 2>&1; echo TEST_EXIT $?; grep -h "MULTI_OK\|passed\|failed\|skipped" gpurun_out/k_tests.log | tail -3
