#!/bin/bash
# One gpurun call: ncu captures of the bandwidth / latency-bound kernels inside a real training step (eager, after warm-up),
# converted to CSV on the box (the .ncu-rep files are too large to bring back) -> gpurun_out/r2_bw_<config>.csv;
# `python tools/bw_summary.py --csv gpurun_out/r2_bw_c2.csv ...` makes the table in profiles/.
set -u
mkdir -p gpurun_out
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy"
R="crop_resize|nms_kernel|rank_count|rank_scatter|avgpool|maxpool|opt_apply|opt_stats|im2col_f32|edgemask|iou_match|sampler_kernel|rpn_decode|softmax_ce|box_classifier|fc_wgrad|fc_fwd|colsum|cast_f32|refine_concat|gather_sampled|detection_targets|rpn_loss|expand_windows|force_match|balanced"
ncu $SEC --clock-control none --profile-from-start off -k regex:"$R" -f -o /tmp/bw_c2 python tools/eager_steps.py --config c2 > gpurun_out/r2_ncu_bw_c2.log 2>&1; echo "bw c2 rc=$?"
ncu -i /tmp/bw_c2.ncu-rep --page raw --csv > gpurun_out/r2_bw_c2.csv 2>/dev/null
ncu $SEC --clock-control none --profile-from-start off -k regex:"psroi" -f -o /tmp/bw_c4 python tools/eager_steps.py --config c4 > gpurun_out/r2_ncu_bw_c4.log 2>&1; echo "bw c4 rc=$?"
ncu -i /tmp/bw_c4.ncu-rep --page raw --csv > gpurun_out/r2_bw_c4.csv 2>/dev/null
ncu $SEC --clock-control none --profile-from-start off -k regex:"dwconv" -f -o /tmp/bw_c1 python tools/eager_steps.py --config c1 > gpurun_out/r2_ncu_bw_c1.log 2>&1; echo "bw c1 rc=$?"
ncu -i /tmp/bw_c1.ncu-rep --page raw --csv > gpurun_out/r2_bw_c1.csv 2>/dev/null
# full set (with source) for the three kernels the verdict names, one launch each
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"crop_resize_fwd|nms_kernel|rank_count" -c 5 -f -o gpurun_out/r2_full_roi_nms python tools/eager_steps.py --config c2 > gpurun_out/r2_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/ | head -30
