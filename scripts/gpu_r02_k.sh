# round 2, call 11 (1 GPU): one shared-memory carve-out for every kernel (so that kernels of the two chains can share an SM) x launch order
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
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/k_$name.json 2> gpurun_out/k_$name.err; echo "$name exit $?"; summ gpurun_out/k_$name.json; }
run maxshared BLISS_X=0
run defaultcarve BLISS_B200_DEFAULT_CARVEOUT=1
run maxshared_order1 BLISS_B200_ORDER=1
run maxshared_order2 BLISS_B200_ORDER=2
run maxshared_order1_prio1 BLISS_B200_ORDER=1 BLISS_B200_STREAM_PRIORITY=1
run maxshared_order2_prio2 BLISS_B200_ORDER=2 BLISS_B200_STREAM_PRIORITY=2
run maxshared_wave512 BLISS_B200_WAVE_SONGS=512
run maxshared_wave512_order1 BLISS_B200_WAVE_SONGS=512 BLISS_B200_ORDER=1
