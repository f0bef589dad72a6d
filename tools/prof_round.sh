# Profiles of the steady-state tracking iteration at 100k kept under profiles/ (run under gpurun; ~3 GPU-minutes):
#   bash tools/prof_round.sh
# tools/prof_iteration.py renders 4 target images and probes the capacity before its eager iterations, so the --set full
# captures skip the matching launches of that set-up (5 forwards) and of the first iteration.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2f_stages_100k.csv python tools/prof_iteration.py 100000 3 > gpurun_out/r2f_stages.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gsd_blend_fwd_chunk|gsd_blend_fwd_finish|gsd_blend_fwd_replay|gsd_tile_sort|gsd_ssim_stats|gsd_ssim_grad|gsd_track_fg_packed" -s 27 -c 7 -o gpurun_out/r2f_top python tools/prof_iteration.py 100000 3 > gpurun_out/r2f_top.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gsd_blend_bwd_chunk|gsd_blend_bwd_prefix|gsd_preprocess_bwd_update|gsd_preprocess_kernel|gsd_bin_" -s 32 -c 7 -o gpurun_out/r2f_bwd python tools/prof_iteration.py 100000 3 > gpurun_out/r2f_bwd.log 2>&1
python tools/graph_timeline.py 100000 20 > gpurun_out/r2f_timeline.txt 2>/dev/null
python tools/step_ablate.py 100000 400 base,no_priors,no_morton,after_fwd,base > gpurun_out/r2f_ablate.txt 2>&1
tail -3 gpurun_out/r2f_ablate.txt
