set -x
python bench.py --steps 3 --warmup 3 --no-sequence --no-train --no-cpu-baseline --single-mode > gpurun_out/r2g_pre.json 2> gpurun_out/r2g_pre.err || exit 1
L=$(python -c "import json;d=json.load(open('gpurun_out/r2g_pre.json'));print(d['gpu_launches']//d['steps'])")
echo launches per step $L
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-sequence --no-train --no-cpu-baseline --single-mode"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3*L)) --launch-count $L --csv --log-file gpurun_out/r2g_launches_step_b6.csv $BENCH > gpurun_out/r2g_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2g_launches_step_b6.csv > gpurun_out/r2g_launches_summary_b6.txt 2>&1
head -30 gpurun_out/r2g_launches_summary_b6.txt
