mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
BLISS_B200_VARIANT=32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stft8192_r64" -c 1 -o gpurun_out/prof_r64 python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --no-cpu-baseline > gpurun_out/ncu_r64.log 2>&1; echo NCU_EXIT $?
