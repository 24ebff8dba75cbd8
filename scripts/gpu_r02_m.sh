# round 2, call 14 (1 GPU): stft8192v2 with the rotating bookkeeping thread + threshold-first pip_track; chroma ring depth 2 / 4
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
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or stages or stft8192v2 or ragged or subslices" > gpurun_out/m_tests.log 2>&1; echo TEST_EXIT $?; tail -6 gpurun_out/m_tests.log | cut -c1-300
BLISS_B200_VARIANT=0 timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/m_v0.json 2> gpurun_out/m_v0.err; echo "default exit $?"; summ gpurun_out/m_v0.json
for k in k5p4 k5p2; do
  BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_$k.so timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/m_$k.json 2> gpurun_out/m_$k.err; echo "$k exit $?"; summ gpurun_out/m_$k.json
done
