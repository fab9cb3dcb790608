#!/bin/bash
# Round profile recipe (GPU box): the bench line, then per workload the ncu launch list of the same command with the DRAM
# byte counters (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; --clock-control none).
# usage: tools/profile_round.sh <tag>     -> gpurun_out/<tag>_*.{json,csv,log}
TAG=${1:-r02}
OUT=gpurun_out
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
for W in cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $W --quick --no-cpu-baseline --steps 3 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err || exit 1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
      --log-file $OUT/${TAG}_launches_$W.csv python bench.py --workload $W --quick --no-cpu-baseline --steps 3 --warmup 3 \
      > $OUT/${TAG}_ncu_$W.log 2>&1
done
