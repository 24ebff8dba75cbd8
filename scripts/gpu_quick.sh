mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/testq.log 2>&1; echo TEST_EXIT $?; tail -3 gpurun_out/testq.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/benchq.json 2> gpurun_out/benchq.err; echo BENCH_EXIT $?
python -c "import json; d=json.load(open('gpurun_out/benchq.json')); print({k: d[k] for k in ('value','ms_per_step','bitwise_reproducible_across_steps','gpu_launches')}); print(d['e2e']); print(d['cpu_baseline'])"
