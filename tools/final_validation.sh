set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/final_bf16.json 2> gpurun_out/final_bf16.err; tail -c 600 gpurun_out/final_bf16.json
python bench.py --dtype tf32 --no-cpu-baseline > gpurun_out/final_tf32.json 2> gpurun_out/final_tf32.err; head -c 300 gpurun_out/final_tf32.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null; head -c 250 gpurun_out/final_ref.json
L=$(python -c "import json;d=json.load(open('gpurun_out/final_bf16.json'));print(d['gpu_launches']//d['steps'])")
echo launches per step $L
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3*L)) --launch-count $L --csv --log-file gpurun_out/r1_launches_step_b4.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
tail -c 300 gpurun_out/final_ncu.log
