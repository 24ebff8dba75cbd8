mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/testq.log 2>&1; echo TEST_EXIT $?; tail -2 gpurun_out/testq.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/benchq.json 2> gpurun_out/benchq.err; echo BENCH_EXIT $?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/benchq_ref.json 2> gpurun_out/benchq_ref.err
timeout 300 python bench_distance.py > gpurun_out/benchq_distance.json 2> gpurun_out/benchq_distance.err; cat gpurun_out/benchq_distance.json
for mb in 128 256 512; do BLISS_B200_TRACE=1 BLISS_B200_CHUNK_MB=$mb timeout 200 python scripts/e2e_probe.py 256 2>&1 | tail -2; done > gpurun_out/e2e_probe3.log 2>&1
cat gpurun_out/e2e_probe3.log
