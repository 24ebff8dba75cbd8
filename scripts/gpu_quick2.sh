mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t2.log 2>&1; echo TEST_EXIT $?; tail -8 gpurun_out/t2.log | cut -c1-200
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % (d.get('e2e') or {}).get('value', 0), 'repro', d.get('bitwise_reproducible_across_steps'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b2_new.json 2> gpurun_out/b2_new.err; echo BENCH_EXIT $?; tail -3 gpurun_out/b2_new.err; summ gpurun_out/b2_new.json
for v in 16 24; do BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b2_var$v.json 2> gpurun_out/b2_var$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/b2_var$v.json; done
BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_k1mb3.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b2_k1mb3.json 2> gpurun_out/b2_k1mb3.err; echo "K1MB3 exit $?"; summ gpurun_out/b2_k1mb3.json
