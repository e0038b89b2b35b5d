set -x
export PYTHONUNBUFFERED=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
T="python scripts/profile_target.py tail 6"
$T > gpurun_out/plain_tail.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tail_fused -s 4 -c 1 -o gpurun_out/r2_tail_p3 $T > gpurun_out/ncu_tail.log 2>&1
CIR_PROFILE_P=2.7 $T > gpurun_out/plain_tail27.log 2>&1 && CIR_PROFILE_P=2.7 ncu --set full --clock-control none --import-source on -k regex:tail_fused -s 4 -c 1 -o gpurun_out/r2_tail_p27 $T > gpurun_out/ncu_tail27.log 2>&1
S="python scripts/profile_target.py search70 4"
$S > gpurun_out/plain_s70.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_search70_launches.csv $S > gpurun_out/ncu_s70.log 2>&1
S="python scripts/profile_target.py search10k 3"
$S > gpurun_out/plain_s10k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 3 -c 1 -o gpurun_out/r2_search10k $S > gpurun_out/ncu_s10k.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv | tail
