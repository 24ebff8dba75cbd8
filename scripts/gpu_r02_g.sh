# round 2, call 7 (1 GPU): packed distance kernel (tests + 100k x 100k A/B), roll-off diagnostic on the config-5 corpus
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "distance or closest or dedup or playlist" > gpurun_out/g_tests.log 2>&1; echo TEST_EXIT $?; tail -15 gpurun_out/g_tests.log | cut -c1-300
for v in 0 65536; do BLISS_B200_VARIANT=$v timeout 300 python bench_distance.py > gpurun_out/g_dist_v$v.json 2> gpurun_out/g_dist_v$v.err; echo "DIST $v exit $?"; cat gpurun_out/g_dist_v$v.json; done
timeout 900 python scripts/diag_rolloff.py 768 > gpurun_out/g_diag.log 2>&1; echo DIAG_EXIT $?; tail -25 gpurun_out/g_diag.log | cut -c1-400
