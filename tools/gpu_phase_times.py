"""Cumulative phase times of the CUDA-graph forward: the graph is cut after each phase (FCVSR_STOP_AFTER) and replayed."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch
    sys.path.insert(0, ROOT)
    from fcvsr_b200 import arch
    from oracle.make_golden import make_clip
    dev = torch.device("cuda:0")
    m = arch.GShiftNet().to(dev).eval()
    m.load_state_dict(arch.seeded_state_dict("full", 0))
    m.compute_dtype = sys.argv[2]
    x = make_clip(1, int(sys.argv[3]), 180, 320).to(dev)
    with torch.no_grad():
        m(x)
        m._engine.use_graph = True
        for _ in range(3):
            m(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / 10:.3f}")
else:
    dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    batch = sys.argv[2] if len(sys.argv) > 2 else "4"
    prev = 0.0
    for stop in ("mgaa_pair", "mgaa", "mffr", "scnet", ""):
        env = dict(os.environ, FCVSR_STOP_AFTER=stop)
        out = subprocess.run([sys.executable, __file__, "--one", dtype, batch], env=env, capture_output=True, text=True)
        try:
            ms = float(out.stdout.strip().splitlines()[-1])
        except Exception:
            print(out.stdout, out.stderr)
            raise
        print(f"through {stop or 'tail (whole forward)':22s}: {ms:7.3f} ms   (+{ms - prev:6.3f})")
        prev = ms
