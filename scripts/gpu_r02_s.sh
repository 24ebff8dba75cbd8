#!/bin/bash
# round 2, call s: the register-resident resampler kernel (4 | down): kernel times, tests, sanitizer, bench line
mkdir -p gpurun_out
cat > /tmp/rs_time.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import bliss_rs_b200 as B
nat = B.native
nat.init(0)
rng = np.random.default_rng(0)
for rate in (44100, 88200, 48000, 96000, 32000, 16000, 8000, 192000, 11025):
    n = int(180 * rate)
    x = rng.standard_normal(n).astype(np.float32)
    for _ in range(2):
        y = nat.resample(x, rate)
    print(rate, y.size)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resample --csv --log-file gpurun_out/s_resample_launches.csv python /tmp/rs_time.py > gpurun_out/s_rs.log 2>&1
echo "NCU_RS exit $?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1
echo "TEST_EXIT $?"; tail -5 gpurun_out/s_tests.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/resample_probe.py > gpurun_out/s_memcheck.log 2>&1; echo MEMCHECK_EXIT $?; grep -E "ERROR SUMMARY|Invalid|probe ok" gpurun_out/s_memcheck.log | head -6
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 --kernel-name kns=resample python scripts/resample_probe.py > gpurun_out/s_racecheck.log 2>&1; echo RACECHECK_EXIT $?; grep -E "RACECHECK SUMMARY|Race reported|hazard|probe ok" gpurun_out/s_racecheck.log | head -6
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "BENCH exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "s16", d["e2e_s16"]["value"])
print("  e2e_cd", json.dumps(d.get("e2e_cd")))
PY
