mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t4.log 2>&1; echo TEST_EXIT $?; tail -6 gpurun_out/t4.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b4.json 2> gpurun_out/b4.err; echo BENCH_EXIT $?; tail -3 gpurun_out/b4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b4.json'))
print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step'])); print('e2e', d['e2e']); print('e2e_s16', d['e2e_s16']); print('cpu', d['cpu_baseline'])
PY
