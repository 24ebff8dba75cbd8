# round-2 closing run (1 GPU): the GPU suite and the bench line on the final build, the ncu capture of the STFT pair kernel
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/y_tests.log 2>&1; echo TEST_EXIT $?; grep -h "passed\|failed" gpurun_out/y_tests.log | tail -2; grep -h "256 x 3-min\|decisions differing: [0-9]* / 256" gpurun_out/y_tests.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; echo BENCH_EXIT $?; tail -3 gpurun_out/y_bench.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/y_ref.json 2> gpurun_out/y_ref.err; echo REF_EXIT $?
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/y_bench.json'))
    print('value %.0f ms/step %.2f e2e %.0f (frac %.3f) s16 %.0f stft frac_read %.3f cpu %.1f parity %.2e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], d['e2e_s16']['value'], d['stft_microbench']['frac_read'], d['cpu_baseline']['value'], d['cpu_baseline']['parity_max_abs_err']))
    print([(k['kernel'], round(k['avg_ms'],2)) for k in d['roofline']['kernels']], d['clocks'])
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stft512_pairs_kernel" -c 1 -o gpurun_out/y_prof_stft512 python bench_stft.py --tracks 128 --resident 128 --warmup 0 > gpurun_out/y_ncu_stft.log 2>&1; echo NCU_STFT_EXIT $?
