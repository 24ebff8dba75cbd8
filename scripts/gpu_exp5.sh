mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:4]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for v in q0 q8 q16 q16c3; do
BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b5_$v.json 2> gpurun_out/b5_$v.err; echo "$v exit $?"; summ gpurun_out/b5_$v.json
done
