mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:4]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
BLISS_B200_VARIANT=32 timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t6.log 2>&1; echo TEST_EXIT $?; tail -8 gpurun_out/t6.log | cut -c1-300
BLISS_B200_VARIANT=32 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b6_r64.json 2> gpurun_out/b6_r64.err; echo BENCH_EXIT $?; tail -3 gpurun_out/b6_r64.err; summ gpurun_out/b6_r64.json
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b6_r16.json 2> gpurun_out/b6_r16.err; summ gpurun_out/b6_r16.json
