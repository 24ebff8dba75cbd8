mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
for v in base k1mb3 k1mb4; do
  if [ $v = base ]; then unset BLISS_B200_SO; else export BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json'))
print('$v', 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:6]))
PY
done
unset BLISS_B200_SO
bash scripts/gpu_e2e_probe.sh
