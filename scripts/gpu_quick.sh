mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/testq.log 2>&1; echo TEST_EXIT $?; tail -3 gpurun_out/testq.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/benchq.json 2> gpurun_out/benchq.err; echo BENCH_EXIT $?
python -c "import json; d=json.load(open('gpurun_out/benchq.json')); print({k: d[k] for k in ('value','ms_per_step','bitwise_reproducible_across_steps','gpu_launches')}); print(d['e2e']); print(d['cpu_baseline'])"
for mb in 32 64 128 256 512; do BLISS_B200_TRACE=1 BLISS_B200_CHUNK_MB=$mb timeout 200 python scripts/e2e_probe.py 256 2>&1 | tail -3; done > gpurun_out/e2e_probe5.log 2>&1
cat gpurun_out/e2e_probe5.log
BLISS_B200_TRACE=1 BLISS_B200_TRACE_CHUNKS=1 BLISS_B200_CHUNK_MB=128 timeout 200 python scripts/e2e_probe.py 256 2>&1 | tail -40 > gpurun_out/e2e_chunks.log; tail -38 gpurun_out/e2e_chunks.log | cut -c1-150
