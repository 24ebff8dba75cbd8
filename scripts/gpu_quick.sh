# quick check: GPU parity tests + one bench line
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/q_tests.log 2>&1; echo TEST_EXIT $?; tail -4 gpurun_out/q_tests.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo BENCH_EXIT $?; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print('value %.0f ms/step %.2f' % (d['value'], d['ms_per_step']), 'repro', d['bitwise_reproducible_across_steps'], 'par', d['cpu_baseline']['parity_max_abs_err'], ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:6]))
PY
