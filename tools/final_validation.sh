set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/final_bf16.json 2> gpurun_out/final_bf16.err; tail -c 600 gpurun_out/final_bf16.json
python bench.py --dtype tf32 --no-cpu-baseline > gpurun_out/final_tf32.json 2> gpurun_out/final_tf32.err; head -c 300 gpurun_out/final_tf32.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null; head -c 250 gpurun_out/final_ref.json
if [ -z "$NO_NCU" ]; then      # the launch list (tools/r2_final_ncu.sh writes the same list)
L=$(python -c "import json;d=json.load(open('gpurun_out/final_bf16.json'));print(d['gpu_launches']//d['steps'])")
echo launches per step $L
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3*L)) --launch-count $L --csv --log-file gpurun_out/final_launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
tail -c 300 gpurun_out/final_ncu.log
fi
if [ -n "$FULL" ]; then
python bench.py --variant S --height 272 --width 480 --batch 2 --no-cpu-baseline > gpurun_out/final_S_272x480.json 2>/dev/null; head -c 160 gpurun_out/final_S_272x480.json; echo
python bench.py --variant S --height 540 --width 960 --batch 1 --steps 10 --no-cpu-baseline > gpurun_out/final_S_540x960.json 2>/dev/null; head -c 160 gpurun_out/final_S_540x960.json; echo
python tools/bench_sequence.py > gpurun_out/final_seq100.json 2>/dev/null; head -c 300 gpurun_out/final_seq100.json; echo
python tools/gpu_phase_times.py 2>&1 | tail -6
fi
