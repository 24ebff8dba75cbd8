mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo TEST_EXIT $?; tail -6 gpurun_out/t3.log | cut -c1-200
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % (d.get('e2e') or {}).get('value', 0), 'repro', d.get('bitwise_reproducible_across_steps'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b3_new.json 2> gpurun_out/b3_new.err; echo BENCH_EXIT $?; tail -3 gpurun_out/b3_new.err; summ gpurun_out/b3_new.json
