#!/bin/bash
# round 2, call u: timedomain beside the chroma STFT -- register caps 56 / 48 / 40 / 32 x shared-memory carve-out preference
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --kernels-only > gpurun_out/u_$name.json 2> gpurun_out/u_$name.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/u_$name.json").read().strip().splitlines()[-1])
print("$name value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]))
PY
}
V=bliss-rs_b200/variants
run td56 X=1
run td56_carve BLISS_B200_MAX_SHARED_CARVEOUT=1
for v in td48 td40 td32; do
  run $v BLISS_B200_SO=$PWD/$V/libbliss_b200_$v.so
  run ${v}_carve BLISS_B200_SO=$PWD/$V/libbliss_b200_$v.so BLISS_B200_MAX_SHARED_CARVEOUT=1
done
