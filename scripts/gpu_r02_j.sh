# round 2, call 10 (1 GPU): scheduling / granularity knobs (launch order of the two chains, items per stft CTA, pairs per pvoc item)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:8]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err; echo "$name exit $?"; summ gpurun_out/j_$name.json; }
run base BLISS_X=0
run order1 BLISS_B200_ORDER=1
run order2 BLISS_B200_ORDER=2
run order1_prio1 BLISS_B200_ORDER=1 BLISS_B200_STREAM_PRIORITY=1
run order1_prio2 BLISS_B200_ORDER=1 BLISS_B200_STREAM_PRIORITY=2
run items8 BLISS_B200_STFT_ITEMS=8
run items16 BLISS_B200_STFT_ITEMS=16
run items2 BLISS_B200_STFT_ITEMS=2
run pairs32 BLISS_B200_PVOC_PAIRS=32
run pairs128 BLISS_B200_PVOC_PAIRS=128
