import sys, torch
sys.path.insert(0, '/root/repo')
from fcvsr_b200 import arch
from fcvsr_b200.engine import Engine
from oracle.make_golden import make_clip
dev = torch.device("cuda:0")
m = arch.GShiftNet().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict("full", 0))
x = make_clip(1, 4, 180, 320).to(dev)
for ms in (True, False, True, False):
    eng = Engine(m, mode="bf16")
    eng.multi_stream = ms
    with torch.no_grad():
        eng.forward(x)
        eng.use_graph = True
        for _ in range(3):
            eng.forward(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"multi_stream={ms}: {e0.elapsed_time(e1) / 20:.3f} ms  {4 * 20 / e0.elapsed_time(e1) * 1e3:.1f} fps")
    del eng
