cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/reh_ref.json 2>gpurun_out/reh_ref.err; echo REF_EXIT $?; python -c "import json; d=json.load(open('gpurun_out/reh_ref.json')); print(d['impl'], d['value'], d['ms_per_step'], d['config'])"
timeout 300 python bench.py > gpurun_out/reh_bench.json 2>gpurun_out/reh_bench.err; echo BENCH_EXIT $?; python -c "
import json; d=json.load(open('gpurun_out/reh_bench.json')); print(sorted(d.keys())); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['roofline']['frac'], d['cpu_baseline']['sample'])"
wc -l gpurun_out/reh_bench.json
