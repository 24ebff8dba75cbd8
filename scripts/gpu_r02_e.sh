# round 2, call 5 (1 GPU): the new bench.py flow end to end (stft_microbench, e2e through measure_e2e), config 5 small
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "BENCH exit $?"; tail -5 gpurun_out/e_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/e_bench.json'))
    print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']))
    print('e2e', json.dumps(d['e2e'])[:700])
    print('e2e_s16', json.dumps(d['e2e_s16'])[:400])
    print('stft', json.dumps(d['stft_microbench'])[:600])
    print('cpu', json.dumps(d['cpu_baseline'])[:900])
    r=d['roofline']; print('roofline', r['kernel'], r['frac'], r['step_read_frac'], [(k['kernel'],round(k['avg_ms'],2)) for k in r['kernels']])
    print('clocks', d['clocks'])
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 900 BLISS_CFG5_SONGS=3000 python bench.py --config 5 --gpus 1 > gpurun_out/e_cfg5_default.json 2> gpurun_out/e_cfg5.err; echo "CFG5 exit $?"; tail -5 gpurun_out/e_cfg5.err | cut -c1-400; cut -c1-2500 gpurun_out/e_cfg5_default.json
