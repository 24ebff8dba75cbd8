# round 2, call 2 (1 GPU): pvoc512v2_kernel on hardware -- targeted parity tests, A/B bench against the round-1 kernel,
# one full ncu capture of the new kernel.
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or stages or silence or ragged or pvoc or implementations or stft512 or white_noise or click" > gpurun_out/b_tests.log 2>&1; echo TEST_EXIT $?; tail -15 gpurun_out/b_tests.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for v in 0 16384; do
  BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/b_v$v.json 2> gpurun_out/b_v$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/b_v$v.json
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pvoc512v2_kernel" -c 1 -o gpurun_out/b_prof_pvoc2 python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --kernels-only > gpurun_out/b_ncu.log 2>&1; echo NCU_EXIT $?
