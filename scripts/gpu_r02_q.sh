#!/bin/bash
# round 2, call q: the resampler (tests, kernel time per song at 44.1 / 48 / 96 kHz, the CD-format e2e leg), full GPU suite
mkdir -p gpurun_out
cat > /tmp/rs_time.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import bliss_rs_b200 as B
nat = B.native
nat.init(0)
rng = np.random.default_rng(0)
for rate in (44100, 48000, 96000, 8000):
    n = int(180 * rate)
    x = rng.standard_normal(n).astype(np.float32)
    for _ in range(3):
        y = nat.resample(x, rate)
    print(rate, y.size)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resample --csv --log-file gpurun_out/q_resample_launches.csv python /tmp/rs_time.py > gpurun_out/q_rs.log 2>&1
echo "NCU_RS exit $?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1
echo "TEST_EXIT $?"; tail -5 gpurun_out/q_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
echo "BENCH exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "s16", d["e2e_s16"]["value"])
print("e2e_cd", json.dumps(d.get("e2e_cd")))
PY
tail -3 gpurun_out/q_bench.err
