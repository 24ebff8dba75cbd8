# round 2, call 18 (1 GPU): small kernels on their own high-priority streams x sub-waves (can the HBM-bound contraction and the
# latency-bound kernels of one wave run under the next wave's FFT kernels?)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'repro', d.get('bitwise_reproducible_across_steps'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/o_$name.json 2> gpurun_out/o_$name.err; echo "$name exit $?"; summ gpurun_out/o_$name.json; tail -2 gpurun_out/o_$name.err | cut -c1-200; }
run base BLISS_X=0
run split BLISS_B200_SPLIT_STREAMS=1
run split_w512 BLISS_B200_SPLIT_STREAMS=1 BLISS_B200_WAVE_SONGS=512
run split_w256 BLISS_B200_SPLIT_STREAMS=1 BLISS_B200_WAVE_SONGS=256
run split_w128 BLISS_B200_SPLIT_STREAMS=1 BLISS_B200_WAVE_SONGS=128
run w256 BLISS_B200_WAVE_SONGS=256
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or ragged or small_workspace or large_batch" > gpurun_out/o_tests.log 2>&1; echo TEST_EXIT $?; tail -3 gpurun_out/o_tests.log | cut -c1-200
BLISS_B200_SPLIT_STREAMS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "golden or ragged or small_workspace or large_batch" > gpurun_out/o_tests_split.log 2>&1; echo TEST_SPLIT_EXIT $?; tail -3 gpurun_out/o_tests_split.log | cut -c1-200
