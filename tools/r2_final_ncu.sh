# Round-2 final profiling pass (run under gpurun on one B200): launch list of one eager bf16 step at the bench's default batch
# (6 windows), `ncu --set full` captures of the dominant kernels inside the model, text summaries into gpurun_out/ (copied to
# profiles/ by hand; the binary reports are not kept).
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 3 --warmup 3 --no-sequence --no-train --no-cpu-baseline --single-mode > gpurun_out/r2f_pre.json 2> gpurun_out/r2f_pre.err || exit 1
L=$(python -c "import json;d=json.load(open('gpurun_out/r2f_pre.json'));print(d['gpu_launches']//d['steps'])")
echo launches per step $L
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-sequence --no-train --no-cpu-baseline --single-mode"
# one eager bf16 step: the model runs 1 (build) + 3 (warm-up) + 1 (timed) forwards before the profiling passes of bench.py itself
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3*L)) --launch-count $L --csv --log-file gpurun_out/r2f_launches_step_b6.csv $BENCH > gpurun_out/r2f_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2f_launches_step_b6.csv > gpurun_out/r2f_launches_summary_b6.txt 2>&1
head -34 gpurun_out/r2f_launches_summary_b6.txt
# SCNet phase: level-batched convolutions (TMA-store epilogue), one-launch ContextBlock, RCB tail, merged down/up, cross-level mix
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|level_mix|rcb_finish_kernel|ctx_block_kernel" --launch-skip 1000 --launch-count 12 -o gpurun_out/r2f_scnet_full -f $BENCH > gpurun_out/r2f_ncu_scnet.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_scnet_full.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel|level_mix|rcb_finish_kernel|ctx_block_kernel --launch-count 12, of: $BENCH (bf16, FCVSR 180x320, 6 windows; SCNetbk phase)" > gpurun_out/r2f_conv_tc_full_summary.txt 2>&1
# MGAA: IAC step with on-chip taps, FFT passes, offset blocks, CorrBlock lookup; tail: conv_last0
ncu --set full --clock-control none --import-source on -k regex:"iac_step_tc_kernel|fft2_|offset_blk|corr_gather|conv3x3_c64_to1" --launch-skip 140 --launch-count 16 -o gpurun_out/r2f_mgaa_full -f $BENCH > gpurun_out/r2f_ncu_mgaa.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_mgaa_full.ncu-rep "ncu --set full: MGAA kernels (IAC step with on-chip taps, FFT passes, offset blocks, CorrBlock lookup), same command" > gpurun_out/r2f_mgaa_full_summary.txt 2>&1
python tools/ncu_table.py gpurun_out/r2f_mgaa_full.ncu-rep > gpurun_out/r2f_mgaa_table.txt 2>&1
python tools/ncu_table.py gpurun_out/r2f_scnet_full.ncu-rep > gpurun_out/r2f_scnet_table.txt 2>&1
ncu --set full --clock-control none -k regex:"conv3x3_c64_to1" --launch-skip 4 --launch-count 1 -o gpurun_out/r2f_last -f python tools/gpu_last_bench.py 6 > gpurun_out/r2f_ncu_last.log 2>&1
python tools/ncu_table.py gpurun_out/r2f_last.ncu-rep > gpurun_out/r2f_last_table.txt 2>&1
rm -f gpurun_out/r2f_*.ncu-rep
python tools/gpu_phase_times.py bf16 4 2>&1 | tail -5 > gpurun_out/r2f_phase_times.txt
python tools/gpu_phase_times.py bf16 6 2>&1 | tail -5 >> gpurun_out/r2f_phase_times.txt
python tools/gpu_conv_bench.py 2>&1 | tail -19 > gpurun_out/r2f_conv_bench.txt
python tools/gpu_iac_bench.py 6 2>&1 | tail -2 > gpurun_out/r2f_iac_bench.txt
python tools/gpu_scnet_bench.py 2>&1 | tail -4 > gpurun_out/r2f_scnet_bench.txt
python tools/gpu_dcn_bench.py 2>&1 | tail -12 > gpurun_out/r2f_dcn_bench.txt
cat gpurun_out/r2f_phase_times.txt
ls -la gpurun_out | tail -20
