#!/bin/bash
# One-shot profiling pass for profiles/ (run under gpurun, ONE GPU):  bash tools/profile_round.sh r02
# 1. plain bench (exit 0 without ncu first)  2. launch list of the timed region  3. --set full capture of every kernel of one
# mapping iteration (+ SASS of the backward: the bulk-copy path)  4. --set full of the dense query / marching cubes / frame
# renderer / tracking kernels  5. microbenchmarks, then the same probes under ncu for their L2 / L1TEX utilisation.
# Raw reports stay in gpurun_out/ (scratch); summaries are written by tools/ncu_summary.py here in the build container.
set -x
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 30 --warmup 5 > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
Q="python bench.py --quick --steps 3 --warmup 3 --no-graph --prefit 2"
timeout 120 $Q > $O/plain.log 2>&1 || exit 1
timeout 300 ncu --nvtx --nvtx-include "usl_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${R}_launches.csv $Q > $O/ncu_launches.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "usl_timed/" \
    -k regex:"field_|composite_|ray_setup|zsample|loss_fwd|fold_|pose_" -c 11 -f -o $O/${R}_full $Q > $O/ncu_full.log 2>&1
ncu -i $O/${R}_full.ncu-rep --page raw --csv > $O/${R}_full_raw.csv 2>/dev/null
ncu -i $O/${R}_full.ncu-rep --page source --csv -k regex:field_bwd2 > $O/${R}_bwd_sass.csv 2>/dev/null
ncu -i $O/${R}_full.ncu-rep --page source --csv -k regex:field_fwd > $O/${R}_fwd_sass.csv 2>/dev/null
[ -n "$SKIP_EXTRAS" ] && exit 0
timeout 200 python tools/profile_targets.py > $O/targets_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none -k regex:"sdf_query_grid|mc_classify|mc_emit|scan_|field_fwd_kernel|composite_fwd|depth_error" -c 14 -f -o $O/${R}_targets \
    python tools/profile_targets.py > $O/ncu_targets.log 2>&1
ncu -i $O/${R}_targets.ncu-rep --page raw --csv > $O/${R}_targets_raw.csv 2>/dev/null
# 4b. the culling and rendering-metrics kernels (tools/bench_isolated.py is also their bench leg)
timeout 300 python tools/bench_isolated.py > $O/${R}_isolated.json 2>&1 && \
timeout 500 ncu --set full --clock-control none -k regex:"mesh_cull|mesh_face_keep|mesh_compact|render_metrics" -c 12 -f -o $O/${R}_cull \
    python tools/bench_isolated.py > $O/ncu_cull.log 2>&1
ncu -i $O/${R}_cull.ncu-rep --page raw --csv > $O/${R}_cull_raw.csv 2>/dev/null
timeout 300 python tools/microbench.py > $O/${R}_microbench.json 2> $O/microbench.err
timeout 300 ncu --metrics gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"bench_" --csv --log-file $O/${R}_microbench_ncu.csv python tools/microbench.py > /dev/null 2> $O/microbench_ncu.err
timeout 200 python tools/microbench_footprint.py > $O/${R}_microbench_footprint.txt 2>&1
exit 0
