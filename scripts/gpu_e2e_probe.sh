mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
for mb in 64 128 192 384 768; do
  BLISS_B200_TRACE=1 BLISS_B200_CHUNK_MB=$mb timeout 200 python scripts/e2e_probe.py 512 2>&1 | grep -E "chunk_mb|trace" | tail -4
done > gpurun_out/e2e_probe.log 2>&1
cat gpurun_out/e2e_probe.log
