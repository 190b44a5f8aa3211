#!/bin/bash
# One-shot profiling pass for profiles/ (run under gpurun, ONE GPU):  bash tools/profile_round.sh r01
# (SKIP_EXTRAS=1: stop after step 3)
# 1. plain bench (exit 0 without ncu first)  2. launch list of the timed region  3. --set full capture of every kernel of
# one mapping iteration  4. the tcgen05 forward variant  5. microbenchmarks.  Raw reports stay in gpurun_out/ (scratch);
# summaries are written by tools/ncu_summary.py here in the build container.
set -x
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 500 python bench.py --steps 30 --warmup 5 > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
Q="python bench.py --quick --steps 3 --warmup 3 --no-graph --prefit 2"
timeout 120 $Q > $O/plain.log 2>&1 || exit 1
timeout 300 ncu --nvtx --nvtx-include "usl_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${R}_launches.csv $Q > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "usl_timed/" \
    -k regex:"field_|composite_|ray_setup|zsample|loss_fwd|fold_|pose_" -c 10 -f -o $O/${R}_full $Q > $O/ncu_full.log 2>&1
ncu -i $O/${R}_full.ncu-rep --page raw --csv > $O/${R}_full_raw.csv 2>/dev/null
[ -n "$SKIP_EXTRAS" ] && exit 0
USL_TCGEN05=1 timeout 120 $Q > $O/plain_tc.log 2>&1 && \
USL_TCGEN05=1 timeout 300 ncu --set full --clock-control none -k regex:field_fwd_tc -c 1 -f -o $O/${R}_tc $Q > $O/ncu_tc.log 2>&1 && \
ncu -i $O/${R}_tc.ncu-rep --page raw --csv > $O/${R}_tc_raw.csv 2>/dev/null
timeout 200 python tools/microbench.py > $O/${R}_microbench.json 2> $O/microbench.err
timeout 200 python tools/microbench_footprint.py > $O/${R}_microbench_footprint.txt 2>&1
exit 0
