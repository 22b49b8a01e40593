set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t7_pytest.log
for sp in 1 2 4; do echo "split $sp" >> gpurun_out/t7_probe.log; SVIN_BACKSUB_SPLIT=$sp timeout 300 python tools/schur_probe.py >> gpurun_out/t7_probe.log 2>&1; done
for mb in 5 6; do echo "minb $mb" >> gpurun_out/t7_probe.log; SVIN_LIN_MINB=$mb timeout 300 python tools/schur_probe.py >> gpurun_out/t7_probe.log 2>&1; done
cat gpurun_out/t7_probe.log
