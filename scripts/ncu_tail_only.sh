set -x
export PYTHONUNBUFFERED=1
T="python scripts/profile_target.py tail 6"
$T > gpurun_out/plain_tail.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tail_fused -s 4 -c 1 -f -o gpurun_out/r2_tail_p3 $T > gpurun_out/ncu_tail.log 2>&1
CIR_PROFILE_P=2.7 $T > gpurun_out/plain_tail27.log 2>&1 && CIR_PROFILE_P=2.7 ncu --set full --clock-control none --import-source on -k regex:tail_fused -s 4 -c 1 -f -o gpurun_out/r2_tail_p27 $T > gpurun_out/ncu_tail27.log 2>&1
ls -la gpurun_out/r2_tail*.ncu-rep
