# round-2 evidence run (1 GPU): the whole GPU suite, the bench line + the reference arm, compute-sanitizer over the new
# kernels, the ncu launch list and one full capture of every kernel of the step.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/z_tests.log 2>&1; echo TEST_EXIT $?; tail -4 gpurun_out/z_tests.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo BENCH_EXIT $?; tail -3 gpurun_out/z_bench.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; echo REF_EXIT $?; cut -c1-300 gpurun_out/z_ref.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/z_bench.json'))
    print('value %.0f ms/step %.2f e2e %.0f (frac %.3f) s16 %.0f stft frac_read %.3f cpu %.1f parity %.2e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], d['e2e_s16']['value'], d['stft_microbench']['frac_read'], d['cpu_baseline']['value'], d['cpu_baseline']['parity_max_abs_err']))
    print([(k['kernel'], round(k['avg_ms'],2)) for k in d['roofline']['kernels']], d['clocks'])
except Exception as e:
    print('bench line unreadable', e)
PY
# compute-sanitizer: memcheck over every kernel (default kernels, then the one-column chroma STFT), racecheck over the kernels written this round
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/repro_probe.py 3 20 1 > gpurun_out/z_memcheck.log 2>&1; echo MEMCHECK_EXIT $?; grep -E "ERROR SUMMARY|Invalid|tempo run0" gpurun_out/z_memcheck.log | head -6
BLISS_B200_VARIANT=131072 timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/repro_probe.py 3 20 1 > gpurun_out/z_memcheck_v3.log 2>&1; echo MEMCHECK_V3_EXIT $?; grep -E "ERROR SUMMARY|Invalid" gpurun_out/z_memcheck_v3.log | head -4
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 --kernel-name kns=pvoc512v2 --kernel-name kns=stft8192v2 --kernel-name kns=finalize python scripts/repro_probe.py 2 20 1 > gpurun_out/z_racecheck.log 2>&1; echo RACECHECK_EXIT $?; grep -E "RACECHECK SUMMARY|Race reported|hazard" gpurun_out/z_racecheck.log | head -6
BLISS_B200_VARIANT=131072 timeout 400 compute-sanitizer --tool racecheck --print-limit 5 --kernel-name kns=stft8192v3 python scripts/repro_probe.py 2 20 1 > gpurun_out/z_racecheck_v3.log 2>&1; echo RACECHECK_V3_EXIT $?; grep -E "RACECHECK SUMMARY|Race reported|hazard" gpurun_out/z_racecheck_v3.log | head -6
# ncu: launch list of a 256-song step (shares), one full capture of every kernel (128 songs)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pvoc512|timedomain|stft8192|tuning|chroma_|peakpick|beattrack|finalize|distance_matrix" -c 60 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --songs-per-gpu 256 --kernels-only > gpurun_out/z_ncu_launch.log 2>&1; echo NCU1_EXIT $?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pvoc512v2_kernel|stft8192v2_kernel|chroma_pipe_kernel|timedomain_kernel|beattrack_kernel|tuning_select_kernel|finalize_kernel|peakpick_kernel|distance_matrix" -c 9 -o gpurun_out/z_prof python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --kernels-only > gpurun_out/z_ncu_full.log 2>&1; echo NCU2_EXIT $?
timeout 300 python bench_distance.py > gpurun_out/z_distance.json 2> gpurun_out/z_distance.err; echo DIST_EXIT $?; cut -c1-400 gpurun_out/z_distance.json
ls -la gpurun_out | grep " z_"
