#!/bin/bash
# round 2, call x: the new chroma-contraction default: GPU suite, bench line, launch list + one full capture of the contraction
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_tests.log 2>&1
echo "TEST_EXIT $?"; tail -3 gpurun_out/x_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
echo "BENCH exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/x_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "s16", d["e2e_s16"]["value"], "cd", d["e2e_cd"]["value"])
for k in d["roofline"]["kernels"]: print("  ", k["kernel"], round(k["avg_ms"], 3))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"chroma_pipe_kernel" -c 1 -o gpurun_out/x_prof_chroma python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --kernels-only > gpurun_out/x_ncu.log 2>&1
echo "NCU exit $?"
ls -la gpurun_out/x_prof_chroma.ncu-rep
