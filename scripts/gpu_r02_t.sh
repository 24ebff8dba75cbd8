#!/bin/bash
# round 2, call t: timedomain with the footprint that fits beside three chroma-STFT CTAs: does it run underneath?
mkdir -p gpurun_out
for order in 0 1 2; do
  BLISS_B200_ORDER=$order timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --kernels-only > gpurun_out/t_order$order.json 2> gpurun_out/t_order$order.err
  echo "ORDER $order exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/t_order$order.json").read().strip().splitlines()[-1])
print("order $order value", d["value"], "ms/step", d["ms_per_step"])
PY
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_tests.log 2>&1
echo "TEST_EXIT $?"; tail -3 gpurun_out/t_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "BENCH exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "s16", d["e2e_s16"]["value"], "cd", d["e2e_cd"]["value"])
for k in d["roofline"]["kernels"]: print("  ", k["kernel"], round(k["avg_ms"], 3))
PY
