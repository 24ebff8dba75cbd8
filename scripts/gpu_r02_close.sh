#!/bin/bash
# round 2, closing call (1 GPU): what the driver runs at round end, on the final build -- GPU suite, smoke(), the bench line with
# its defaults, the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo "TEST_EXIT $?"; tail -2 gpurun_out/c_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0
timeout 600 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "BENCH exit $? in ${SECONDS}s"
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; echo "REF exit $? in ${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/c_bench_ref.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "s16", d["e2e_s16"]["value"], "cd", d["e2e_cd"]["value"], "launches", d["gpu_launches"])
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "reference arm", r["value"], r["cpu_baseline"]["cores"])
print("clocks", d["clocks"])
PY
