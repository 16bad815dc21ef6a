#!/bin/bash
# One profiling pass for profiles/<tag>_*: launch lists of a bench step and of an FGR batch, and `ncu --set full`
# captures (with source) of the dominant kernels, exported as raw-page CSVs.  Run on the GPU box AFTER the same
# commands exited 0 without ncu:  gpurun -- 'tools/profile_round.sh r2_02'
# Writes gpurun_out/<tag>_*.{csv,ncu-rep}; copy the CSVs to profiles/ (the .ncu-rep files stay in scratch).
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"

# 1) launch lists (cold-cache, serialised per-launch times: shares of the step, not absolutes)
$NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_bench_under_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $out/${tag}_fgr_batch_launches.csv \
    python tools/fgr_bench.py fixed 2 > $out/${tag}_fgr_under_ncu.log 2>&1

# 2) full captures
REPS=2 $NCU --set full --import-source on -k regex:gicp_persistent -s 1 -c 1 -o $out/${tag}_gicp_persistent \
    python tools/quick_bench.py > $out/${tag}_cap_gicp.log 2>&1
$NCU --set full --import-source on -k "regex:feat_match|feat_resolve|fpfh_kernel|spfh_kernel|fgr_tuple|fgr_gnc|fgr_prepare|feat_pack" \
    -s 12 -c 14 -o $out/${tag}_fgr_kernels python tools/fgr_bench.py fixed 2 > $out/${tag}_cap_fgr.log 2>&1
REPS=2 $NCU --set full --import-source on -k "regex:knn_wq|knn_rerank_kernel|normals_cov" -s 4 -c 4 -o $out/${tag}_knn_normals \
    python tools/quick_bench.py > $out/${tag}_cap_knn.log 2>&1

# 3) raw pages
for n in gicp_persistent fgr_kernels knn_normals; do
    [ -f $out/${tag}_$n.ncu-rep ] && ncu -i $out/${tag}_$n.ncu-rep --page raw --csv > $out/${tag}_${n}_ncu_raw.csv 2>/dev/null
done
[ -f $out/${tag}_gicp_persistent.ncu-rep ] && ncu -i $out/${tag}_gicp_persistent.ncu-rep --page details --csv > $out/${tag}_gicp_persistent_ncu_details.csv 2>/dev/null
ls -la $out | grep ${tag}_
