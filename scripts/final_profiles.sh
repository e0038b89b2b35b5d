#!/bin/bash
# One GPU box: bench line, extra configs, ncu launch list and the --set full captures (tag = $1).  Every profiled
# command is first run to completion without ncu.
tag=${1:-r1f}
out=gpurun_out
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
python scripts/bench_extra.py > $out/extra_$tag.jsonl 2> $out/extra_$tag.err || tail -3 $out/extra_$tag.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/plain_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
cap() {  # name, kernel regex, skip, target args...
  local name=$1 k=$2 s=$3; shift 3
  python scripts/profile_target.py "$@" > $out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o $out/prof_${name}_$tag \
      python scripts/profile_target.py "$@" > $out/p_$name.log 2>&1
  tail -1 $out/p_$name.log
}
cap tail tail_fused 1 tail 3
cap search10k search_kernel 3 search10k 2
cap prepass10k search_kernel 2 search10k 2
cap select10k topk_select_warp 1 search10k 2
cap kth10k row_kth 1 search10k 2
cap search70 search_kernel 3 search70 2
cap regions region_pool 1 regions 3
ls -la $out/*_$tag*
