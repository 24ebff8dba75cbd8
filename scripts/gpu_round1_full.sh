mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; free -g | head -2
timeout 600 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/test14.log 2>&1; echo TEST_EXIT $?
tail -3 gpurun_out/test14.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo BENCH_EXIT $?
tail -c 3000 gpurun_out/bench14.json; tail -5 gpurun_out/bench14.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench14_ref.json 2> gpurun_out/bench14_ref.err; echo REF_EXIT $?
cat gpurun_out/bench14_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pvoc512|timedomain|stft8192|tuning_kernel|chroma_kernel|peakpick|beattrack|finalize|distance_matrix" -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --songs-per-gpu 256 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo NCU1_EXIT $?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pvoc512_kernel|stft8192_kernel|chroma_kernel|timedomain_kernel|beattrack_kernel|tuning_kernel|finalize_kernel|peakpick_kernel|distance_matrix" -c 9 -o gpurun_out/prof_r01 python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo NCU2_EXIT $?
timeout 300 python bench_stft.py --tracks 4000 --resident 1000 > gpurun_out/bench14_stft.json 2> gpurun_out/bench14_stft.err; echo STFT_EXIT $?; cat gpurun_out/bench14_stft.json; tail -3 gpurun_out/bench14_stft.err
timeout 300 python bench_distance.py > gpurun_out/bench14_distance.json 2> gpurun_out/bench14_distance.err; echo DIST_EXIT $?; cat gpurun_out/bench14_distance.json; tail -3 gpurun_out/bench14_distance.err
BLISS_B200_TRACE=1 timeout 200 python scripts/e2e_probe.py 256 > gpurun_out/e2e_probe2.log 2>&1; tail -8 gpurun_out/e2e_probe2.log
python scripts/h2d_bw.py > gpurun_out/h2d_bw.json 2>&1; cat gpurun_out/h2d_bw.json
ls -la gpurun_out
