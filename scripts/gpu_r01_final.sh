# round-1 evidence run (1 GPU): parity tests, bench + reference arm, A/B against the previous kernels, ncu launch
# list + full capture, STFT and distance micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; free -g | head -2
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/f_tests.log 2>&1; echo TEST_EXIT $?; tail -4 gpurun_out/f_tests.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo BENCH_EXIT $?; tail -3 gpurun_out/f_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo REF_EXIT $?; cat gpurun_out/f_ref.json | cut -c1-600
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % (d.get('e2e') or {}).get('value', 0), 'repro', d.get('bitwise_reproducible_across_steps'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
summ gpurun_out/f_bench.json
BLISS_B200_VARIANT=31 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_var31.json 2> gpurun_out/f_var31.err; echo "VARIANT 31 exit $?"; summ gpurun_out/f_var31.json
BLISS_B200_VARIANT=31 BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_nopacked.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_var31_nopacked.json 2> gpurun_out/f_var31_nopacked.err; echo "VARIANT 31 + scalar FP exit $?"; summ gpurun_out/f_var31_nopacked.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pvoc512|timedomain|stft8192|tuning|chroma_|peakpick|beattrack|finalize|distance_matrix" -c 60 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --songs-per-gpu 256 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1; echo NCU1_EXIT $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pvoc512_kernel|stft8192_kernel|chroma_pipe_kernel|timedomain_kernel|beattrack_kernel|tuning_select_kernel|finalize_kernel|peakpick_kernel|distance_matrix" -c 9 -o gpurun_out/f_prof python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1; echo NCU2_EXIT $?
timeout 200 python bench_stft.py --tracks 4000 --resident 1000 > gpurun_out/f_stft.json 2> gpurun_out/f_stft.err; echo STFT_EXIT $?; cat gpurun_out/f_stft.json
timeout 300 python bench_distance.py > gpurun_out/f_distance.json 2> gpurun_out/f_distance.err; echo DIST_EXIT $?; cat gpurun_out/f_distance.json | cut -c1-500
ls -la gpurun_out | grep " f_"
