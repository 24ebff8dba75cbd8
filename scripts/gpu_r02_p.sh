# round 2, call 19 (1 GPU): split streams x ONE carve-out for all kernels x sub-waves (the two conditions for co-residency together)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/p_$name.json 2> gpurun_out/p_$name.err; echo "$name exit $?"; summ gpurun_out/p_$name.json; tail -2 gpurun_out/p_$name.err | cut -c1-200; }
run base BLISS_X=0
run carve_split_w512 BLISS_B200_MAX_SHARED_CARVEOUT=1 BLISS_B200_SPLIT_STREAMS=1 BLISS_B200_WAVE_SONGS=512
run carve_split_w256 BLISS_B200_MAX_SHARED_CARVEOUT=1 BLISS_B200_SPLIT_STREAMS=1 BLISS_B200_WAVE_SONGS=256
run carve_split BLISS_B200_MAX_SHARED_CARVEOUT=1 BLISS_B200_SPLIT_STREAMS=1
