# round-2 opener (1 GPU): GPU tests (incl. the experimental kernel cuts written blind at the end of round 1),
# then the same bench line once per BLISS_B200_VARIANT mask so that every cut is A/B-timed on one box, then the
# STFT micro-benchmark with and without the hop-256 pair kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r02_ab.sh'      (19 bench lines of about a minute each + tests + two ncu captures)
# masks: 64 stft8192 product twiddles | 128 stft8192 synthesised window | 512 pvoc512 product twiddles |
#        1024 pvoc512 pair descriptors + MUFU-only magnitudes | 2048 pvoc512 conflict-free tile padding |
#        4096 stft8192 conflict-free buffer layout | 256 STFT micro-benchmark pair kernel
#        8192 stft8192 aligned loads for odd-start frames
#        32 the radix-64 stft8192 kernel (measured in round 1), 160 / 8224 / 8352 = it with the window / odd-start / both load cuts
#        (12480 = all stft8192 cuts, 3584 = all pvoc512 cuts, 16064 = everything)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/ab_tests.log 2>&1; echo TEST_EXIT $?; tail -6 gpurun_out/ab_tests.log | cut -c1-300
grep -h "bit-identical" gpurun_out/ab_tests.log | head -3
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'), 'par', (d.get('cpu_baseline') or {}).get('parity_max_abs_err'), ' '.join('%s=%.2f' % (k['kernel'][:8], k['avg_ms']) for k in d['roofline']['kernels'][:7]))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for v in 0 64 128 4096 8192 12480 32 160 8224 8352 512 1024 2048 3584 16064; do
  extra="--no-cpu-baseline"; [ $v = 0 ] && extra=""; [ $v = 16064 ] && extra=""   # parity against the oracle for the default and for everything on
  BLISS_B200_VARIANT=$v timeout 400 python bench.py --steps 5 --warmup 3 $extra > gpurun_out/ab_v$v.json 2> gpurun_out/ab_v$v.err; echo "VARIANT $v exit $?"; summ gpurun_out/ab_v$v.json
done
# stream priorities between the two chains of a wave (api.cu, BLISS_B200_STREAM_PRIORITY): 1 = tempo / timbral chain
# first, 2 = chroma chain first
for pr in 1 2; do
  BLISS_B200_STREAM_PRIORITY=$pr timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_prio$pr.json 2> gpurun_out/ab_prio$pr.err; echo "PRIORITY $pr exit $?"; summ gpurun_out/ab_prio$pr.json
done
# sub-waves of a resident batch (BLISS_B200_WAVE_SONGS): the HBM-bound contraction of one sub-wave under the SM-bound
# FFT kernels of the next (measured flat with the first-half kernels; the contraction has changed since)
for ws in 512 256; do
  BLISS_B200_WAVE_SONGS=$ws timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_wave$ws.json 2> gpurun_out/ab_wave$ws.err; echo "WAVE_SONGS $ws exit $?"; summ gpurun_out/ab_wave$ws.json
done
# magnitude-spill rows on 128-byte lines: a separate build (python scripts/build_variants.py BEFORE the gpurun call)
if [ -f bliss-rs_b200/variants/libbliss_b200_stride4128.so ]; then
  BLISS_B200_SO=$PWD/bliss-rs_b200/variants/libbliss_b200_stride4128.so timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_stride4128.json 2> gpurun_out/ab_stride4128.err; echo "STRIDE 4128 exit $?"; summ gpurun_out/ab_stride4128.json
fi
for v in 0 256; do
  BLISS_B200_VARIANT=$v timeout 300 python bench_stft.py --tracks 4000 --resident 1000 --cufft > gpurun_out/ab_stft_v$v.json 2> gpurun_out/ab_stft_v$v.err; echo "STFT VARIANT $v exit $?"; cut -c1-400 gpurun_out/ab_stft_v$v.json
done
# one full ncu capture of the two FFT kernels with every cut on (compare with profiles/ncu_r01b_full_128songs.md:
# data-pipe wavefronts, bank conflicts, issue slots) and of the STFT pair kernel
BLISS_B200_VARIANT=16064 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pvoc512_kernel|stft8192_kernel" -c 2 -o gpurun_out/ab_prof_v16064 python bench.py --steps 1 --warmup 0 --songs-per-gpu 128 --no-cpu-baseline > gpurun_out/ab_ncu_v16064.log 2>&1; echo NCU_EXIT $?
BLISS_B200_VARIANT=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stft512_pairs_kernel" -c 1 -o gpurun_out/ab_prof_stft_v256 python bench_stft.py --tracks 128 --resident 128 --warmup 0 > gpurun_out/ab_ncu_stft_v256.log 2>&1; echo NCU_STFT_EXIT $?
ls -la gpurun_out | grep " ab_"
python scripts/ab_report.py gpurun_out
