# round 2, call 8 (1 GPU): stft8192v3_kernel + the sequential-mean finalize: tests, A/B against v2, ncu
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
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or stages or stft8192v2 or ragged or subslices or 2_pow_24 or silence or three_minute or corpus" > gpurun_out/i_tests_quick.log 2>&1; echo QUICK_TEST_EXIT $?; tail -12 gpurun_out/i_tests_quick.log | cut -c1-300
for v in 0 131072; do
  BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/i_v$v.json 2> gpurun_out/i_v$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/i_v$v.json; tail -3 gpurun_out/i_v$v.err | cut -c1-300
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"stft8192v3_kernel" -c 1 -o gpurun_out/i_prof_stft3 python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --kernels-only > gpurun_out/i_ncu.log 2>&1; echo NCU_EXIT $?
timeout 900 python scripts/diag_rolloff.py 256 > gpurun_out/i_diag.log 2>&1; echo DIAG_EXIT $?; grep "max err per feature\|worst songs" gpurun_out/i_diag.log | cut -c1-300
