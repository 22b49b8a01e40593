set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_dense_eval|k_dense_gram|k_dense_solve_smem|k_step_dense|k_step_lm|k_backsub|k_linearize" -s 14 -c 7 -o gpurun_out/t5_prof -f python tools/schur_probe.py --windows 256 --solves 1 --no-prof > gpurun_out/t5_ncu.log 2>&1; echo "ncu rc=$?"
SVIN_BA_NO_FORK=1 timeout 300 python tools/schur_probe.py --no-prof --solves 3 >> gpurun_out/t5_probe.log 2>&1
timeout 300 python tools/schur_probe.py --no-prof --solves 3 >> gpurun_out/t5_probe.log 2>&1
timeout 300 python tools/schur_probe.py --no-prof --solves 3 --windows 296 >> gpurun_out/t5_probe.log 2>&1
cat gpurun_out/t5_probe.log
