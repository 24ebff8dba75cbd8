mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
for v in 0 8; do echo "== VARIANT $v: 1024 x 30 s"; BLISS_B200_VARIANT=$v timeout 120 python scripts/repro_probe.py 1024 30 3 2>&1 | tail -4; done
for v in 0 8; do echo "== VARIANT $v: 4 x 30 s"; BLISS_B200_VARIANT=$v timeout 120 python scripts/repro_probe.py 4 30 4 2>&1 | tail -5; done
echo "== racecheck beattrack (new ACF)"
timeout 400 compute-sanitizer --tool racecheck --kernel-name kns=beattrack --print-limit 20 python scripts/repro_probe.py 2 30 1 > gpurun_out/race_new.log 2>&1; tail -40 gpurun_out/race_new.log
echo "== racecheck beattrack (old ACF)"
BLISS_B200_VARIANT=8 timeout 400 compute-sanitizer --tool racecheck --kernel-name kns=beattrack --print-limit 20 python scripts/repro_probe.py 2 30 1 > gpurun_out/race_old.log 2>&1; tail -25 gpurun_out/race_old.log
echo "== initcheck / memcheck beattrack + peakpick + pvoc"
timeout 400 compute-sanitizer --tool initcheck --kernel-name kns=beattrack --kernel-name kns=peakpick --kernel-name kns=pvoc512 --print-limit 10 python scripts/repro_probe.py 2 30 1 > gpurun_out/init_new.log 2>&1; tail -25 gpurun_out/init_new.log
