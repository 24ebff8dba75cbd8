# round-2 opener (1 GPU), trimmed: pipe-rate micro-benchmarks, then the same bench line (kernels only) once per
# BLISS_B200_VARIANT mask so that every experimental cut of round 1 is A/B-timed on one box.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r02_ab.sh'
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
./scripts/ubench > gpurun_out/ubench.txt 2>&1; echo UBENCH_EXIT $?; cat gpurun_out/ubench.txt
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for v in 0 64 128 4096 8192 12480 8352 512 1024 2048 3584 16064; do
  BLISS_B200_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/ab_v$v.json 2> gpurun_out/ab_v$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/ab_v$v.json
done
for pr in 1 2; do
  BLISS_B200_STREAM_PRIORITY=$pr timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/ab_prio$pr.json 2> gpurun_out/ab_prio$pr.err; echo "PRIORITY $pr exit $?"; summ gpurun_out/ab_prio$pr.json
done
BLISS_B200_WAVE_SONGS=256 timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/ab_wave256.json 2> gpurun_out/ab_wave256.err; echo "WAVE_SONGS 256 exit $?"; summ gpurun_out/ab_wave256.json
if [ -f bliss-rs_b200/variants/libbliss_b200_stride4128.so ]; then
  BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_stride4128.so timeout 300 python bench.py --steps 4 --warmup 3 --kernels-only > gpurun_out/ab_stride4128.json 2> gpurun_out/ab_stride4128.err; echo "STRIDE 4128 exit $?"; summ gpurun_out/ab_stride4128.json
fi
for v in 0 256; do
  BLISS_B200_VARIANT=$v timeout 300 python bench_stft.py --tracks 4000 --resident 1000 --cufft > gpurun_out/ab_stft_v$v.json 2> gpurun_out/ab_stft_v$v.err; echo "STFT VARIANT $v exit $?"; cut -c1-400 gpurun_out/ab_stft_v$v.json
done
python scripts/ab_report.py gpurun_out
