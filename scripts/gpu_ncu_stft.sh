# ncu capture of the STFT-only micro-benchmark kernel (BASELINE.json configs[2]: 512-point, hop 256, 257 magnitudes)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 40 ncu --set full --clock-control none -k regex:pvoc512_kernel -c 1 -o gpurun_out/prof_stft512 python bench_stft.py --tracks 128 --resident 128 --warmup 0 > gpurun_out/ncu_stft512.log 2>&1; echo NCU_EXIT $?
