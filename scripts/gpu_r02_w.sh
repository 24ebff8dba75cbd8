#!/bin/bash
# round 2, call w: chroma contraction, swizzled 64-byte stage rows at four CTAs per SM (128 registers)
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/w_$name.json 2> gpurun_out/w_$name.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/w_$name.json").read().strip().splitlines()[-1])
k = {x["kernel"]: x["avg_ms"] for x in d["roofline"]["kernels"]}
print("$name value %.0f ms/step %.3f chroma %.3f stft %.3f pvoc %.3f parity %s" % (d["value"], d["ms_per_step"], k["chroma_kernel"], k["stft8192_kernel"], k["pvoc512_kernel"], d["cpu_baseline"].get("parity_max_abs_err")))
PY
}
V=bliss-rs_b200/variants
run k5p3sw BLISS_B200_SO=$PWD/$V/libbliss_b200_k5p3sw.so
run k5p3sw4 BLISS_B200_SO=$PWD/$V/libbliss_b200_k5p3sw4.so
run k5p2sw4 BLISS_B200_SO=$PWD/$V/libbliss_b200_k5p2sw4.so
