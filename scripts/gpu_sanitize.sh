# compute-sanitizer over a tiny batch (3 songs x 20 s): memcheck on every kernel, racecheck on the kernels that
# were rewritten this round.  Also a plain run first (sanity of the build).
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 60 python scripts/repro_probe.py 64 30 2 2>&1 | tail -3
timeout 70 compute-sanitizer --tool memcheck --print-limit 5 python scripts/repro_probe.py 3 20 1 > gpurun_out/memcheck.log 2>&1; echo MEMCHECK_EXIT $?; grep -E "ERROR SUMMARY|Invalid|tempo run0" gpurun_out/memcheck.log | head -6
timeout 100 compute-sanitizer --tool racecheck --print-limit 5 --kernel-name kns=chroma_pipe --kernel-name kns=tuning_select --kernel-name kns=stft8192 python scripts/repro_probe.py 2 20 1 > gpurun_out/racecheck.log 2>&1; echo RACECHECK_EXIT $?; grep -E "RACECHECK SUMMARY|Race reported|tempo run0" gpurun_out/racecheck.log | head -6
