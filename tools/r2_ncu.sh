# Round-2 profiling pass (run under gpurun on one B200): launch list of one eager step, `ncu --set full` captures of the
# dominant kernels inside the model, text summaries into gpurun_out/ (copied to profiles/ by hand).
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 3 --warmup 3 --no-sequence --no-train --no-cpu-baseline --single-mode > gpurun_out/r2_pre.json 2> gpurun_out/r2_pre.err || exit 1
L=$(python -c "import json;d=json.load(open('gpurun_out/r2_pre.json'));print(d['gpu_launches']//d['steps'])")
echo launches per step $L
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-sequence --no-train --no-cpu-baseline --single-mode"
# one eager bf16 step: the model runs 1 (build) + 3 (warm-up) + 1 (timed) forwards before the profiling passes of bench.py itself
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3*L)) --launch-count $L --csv --log-file gpurun_out/r2_launches_step_b4.csv $BENCH > gpurun_out/r2_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_step_b4.csv > gpurun_out/r2_launches_summary_b4.txt 2>&1
head -30 gpurun_out/r2_launches_summary_b4.txt
# full captures: SCNet phase (level-batched convolutions and helpers); skip the MGAA / MFFR launches of the 4th forward
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|level_mix_kernel|rcb_finish_kernel|ctx_partial_kernel|ctx_finalize_kernel" --launch-skip 1190 --launch-count 14 -o gpurun_out/r2_scnet_full -f $BENCH > gpurun_out/r2_ncu_scnet.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_scnet_full.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel|level_mix_kernel|rcb_finish_kernel|ctx_* --launch-count 14, of: $BENCH (bf16, FCVSR 180x320, 4 windows; SCNetbk phase, level-batched launches)" > gpurun_out/r2_conv_tc_full_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"iac_step_kernel|fft2_|offset_blk|corr_gather" --launch-skip 159 --launch-count 14 -o gpurun_out/r2_mgaa_full -f $BENCH > gpurun_out/r2_ncu_mgaa.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_mgaa_full.ncu-rep "ncu --set full: MGAA kernels (IAC step, FFT passes, offset blocks, CorrBlock lookup), same command" > gpurun_out/r2_mgaa_full_summary.txt 2>&1
# training step: tcgen05 weight gradient, data gradient, flow_warp / SAC adjoints
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|conv_wgrad_kernel|flow_warp_bwd|sac_bwd|colsum_partial" --launch-skip 2500 --launch-count 10 -o gpurun_out/r2_train_full -f python tools/gpu_train_once.py > gpurun_out/r2_ncu_train.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_train_full.ncu-rep "ncu --set full: training-step kernels (tcgen05 wgrad, adjoints), python tools/gpu_train_once.py (FCVSR, batch 8 of 7x64x64)" > gpurun_out/r2_train_full_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -15
