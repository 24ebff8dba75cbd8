# round 2, call 13 (1 GPU): balanced autocorrelation of beattrack_kernel: tempo tests + A/B against the four-lags-per-thread cut (bit 262144)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:8]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or click or implementations or silence or corpus or three_minute or large_batch" > gpurun_out/l_tests.log 2>&1; echo TEST_EXIT $?; tail -6 gpurun_out/l_tests.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import bliss_rs_b200 as B
from bliss_rs_b200 import synth
nat = B.native; nat.init(0)
songs = [synth.gen_track(5, i, 22050 * 100 + 977 * i, device="cuda:0").cpu().numpy() for i in range(48)]
_, f0 = nat.analyze_batch(songs, 2)
for mask in (262144, 8):
    nat.set_variant(mask); _, f = nat.analyze_batch(songs, 2); nat.set_variant(0)
    print("mask", mask, "bit-identical to the balanced autocorrelation:", bool(np.array_equal(f, f0)))
PY
for v in 0 262144; do
  BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/l_v$v.json 2> gpurun_out/l_v$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/l_v$v.json
done
