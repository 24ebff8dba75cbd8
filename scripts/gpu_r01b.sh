# one gpurun call: parity tests on the current kernels, A/B against the previous kernels (BLISS_B200_VARIANT,
# scalar-FP build), bisect on failure, ncu launch list + full capture.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_new.log 2>&1; T=$?; echo TEST_EXIT $T; tail -15 gpurun_out/t_new.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b_new.json 2> gpurun_out/b_new.err; echo BENCH_NEW_EXIT $?; tail -3 gpurun_out/b_new.err
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
summ gpurun_out/b_new.json
for v in 15; do
  BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_var$v.json 2> gpurun_out/b_var$v.err; echo "VARIANT $v exit $?"
  summ gpurun_out/b_var$v.json
done
BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_nopacked.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_nopacked.json 2> gpurun_out/b_nopacked.err; echo "NOPACKED exit $?"
summ gpurun_out/b_nopacked.json
if [ $T -ne 0 ]; then
  for v in 1 2 4 8; do
    BLISS_B200_VARIANT=$v timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/t_var$v.log 2>&1; echo "TEST VARIANT $v exit $?"; tail -4 gpurun_out/t_var$v.log
  done
  BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_nopacked.so timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/t_nopacked.log 2>&1; echo "TEST NOPACKED exit $?"; tail -4 gpurun_out/t_nopacked.log
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pvoc512|timedomain|stft8192|tuning|chroma_|peakpick|beattrack|finalize|distance_matrix" -c 60 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --songs-per-gpu 256 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo NCU1_EXIT $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pvoc512_kernel|stft8192_kernel|chroma_pipe_kernel|timedomain_kernel|beattrack_kernel|tuning_select_kernel|finalize_kernel|peakpick_kernel|distance_matrix" -c 9 -o gpurun_out/prof_r01b python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo NCU2_EXIT $?
timeout 200 python bench_stft.py --tracks 4000 --resident 1000 > gpurun_out/b_stft.json 2> gpurun_out/b_stft.err; echo STFT_EXIT $?; cat gpurun_out/b_stft.json
ls -la gpurun_out
