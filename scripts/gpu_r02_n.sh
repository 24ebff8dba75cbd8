# round 2, call 15 (1 GPU): the two stft8192v2 tweaks separately (rotating bookkeeping thread, threshold-first pip_track), longer runs
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
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err; echo "$name exit $?"; summ gpurun_out/n_$name.json; }
run base BLISS_X=0
run s2rot BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_s2rot.so
run s2hot BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_s2hot.so
run items32 BLISS_B200_STFT_ITEMS=32
run pairs256 BLISS_B200_PVOC_PAIRS=256
