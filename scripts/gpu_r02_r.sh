#!/bin/bash
# round 2, call r: resampler kernels v2 (decimation / shared-table) + the conversion stream: tests, kernel times, e2e legs A/B
mkdir -p gpurun_out
cat > /tmp/rs_time.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import bliss_rs_b200 as B
nat = B.native
nat.init(0)
rng = np.random.default_rng(0)
for variant in (0, 524288):
    nat.set_variant(variant)
    for rate in (44100, 88200, 48000, 96000, 32000, 8000):
        n = int(180 * rate)
        x = rng.standard_normal(n).astype(np.float32)
        for _ in range(2):
            y = nat.resample(x, rate)
        print(variant, rate, y.size)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resample --csv --log-file gpurun_out/r_resample_launches.csv python /tmp/rs_time.py > gpurun_out/r_rs.log 2>&1
echo "NCU_RS exit $?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.log 2>&1
echo "TEST_EXIT $?"; tail -5 gpurun_out/r_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "BENCH exit $?"
BLISS_B200_CONV_ON_COPY_STREAM=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r_bench_convoncopy.json 2> gpurun_out/r_bench2.err
echo "BENCH2 exit $?"
python - <<'PY'
import json
for f in ("r_bench", "r_bench_convoncopy"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "s16", d["e2e_s16"]["value"])
    print("  e2e_cd", json.dumps(d.get("e2e_cd")))
PY
tail -3 gpurun_out/r_bench.err
