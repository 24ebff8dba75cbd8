# round 2 (8 GPUs, one box): the bench line at N=8 (incl. e2e through ONE multi-device call), BASELINE configs[3]
# (100 k tracks) and configs[4] (Zipf corpus, 20 k songs), the multi-device / gather tests.
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8; nproc; free -g | head -2; nvidia-smi topo -m 2>/dev/null | head -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/h_bench8.json 2> gpurun_out/h_bench8.err; echo "BENCH8 exit $?"; tail -4 gpurun_out/h_bench8.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/h_bench8.json'))
    print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']))
    print('e2e', json.dumps(d['e2e'])[:900])
    print('e2e_s16', json.dumps(d['e2e_s16'])[:300])
    print('gather', d['gather'], 'per_rank', d['per_rank'], 'clocks', d['clocks'])
except Exception as e:
    print('bench line unreadable', e)
PY
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR4 --master-port 29524 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/h_bench4.json 2> gpurun_out/h_bench4.err; echo "BENCH4 exit $?"; python -c "
import json
d=json.load(open('gpurun_out/h_bench4.json')); print('N=4 value %.0f e2e %.0f frac %.3f ceiling %.1f GB/s s16 %.0f' % (d['value'], d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], d['e2e']['h2d_ceiling_gbs'], d['e2e_s16']['value']))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 $TR --master-port 29522 bench.py --config 4 --gpus 8 > gpurun_out/h_cfg4.json 2> gpurun_out/h_cfg4.err; echo "CFG4 exit $?"; tail -4 gpurun_out/h_cfg4.err | cut -c1-300; cut -c1-1500 gpurun_out/h_cfg4.json
timeout 1200 python bench.py --config 5 --gpus 8 > gpurun_out/h_cfg5.json 2> gpurun_out/h_cfg5.err; echo "CFG5 exit $?"; tail -4 gpurun_out/h_cfg5.err | cut -c1-300; cut -c1-3000 gpurun_out/h_cfg5.json
timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_gather.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/h_tests.log 2>&1; echo TEST_EXIT $?; grep -h "MULTI_OK\|passed\|failed\|skipped" gpurun_out/h_tests.log | tail -4
timeout 300 $TR --master-port 29523 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/h_ref8.json 2> gpurun_out/h_ref8.err; echo "REF8 exit $?"; cut -c1-600 gpurun_out/h_ref8.json
