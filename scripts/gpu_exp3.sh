mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for v in k5contig k5s4; do
BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b3_$v.json 2> gpurun_out/b3_$v.err; echo "$v exit $?"; summ gpurun_out/b3_$v.json
done
