# Session profiling pass: standalone IAC timing + ncu --set full captures of iac_step_tc, ctx_partial, offset_blk_conv inside one eager bench step
set -x
cd "$(dirname "$0")/.."
python tools/gpu_iac_bench.py 6 2>&1 | tail -3
python tools/gpu_iac_bench.py 4 2>&1 | tail -3
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-sequence --no-train --no-cpu-baseline --single-mode"
ncu --set full --clock-control none --import-source on -k regex:"iac_step_tc_kernel" --launch-skip 78 --launch-count 2 -o gpurun_out/s2_iac -f $BENCH > gpurun_out/s2_ncu_iac.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ctx_partial_kernel|offset_blk_conv_kernel|ctx_finalize" --launch-skip 150 --launch-count 6 -o gpurun_out/s2_misc -f $BENCH > gpurun_out/s2_ncu_misc.log 2>&1
ls -la gpurun_out | tail
