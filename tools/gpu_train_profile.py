"""Where does the config-4 training step spend its time?  torch.profiler over two steps: device time per kernel (top 25),
total device-busy time against the wall time of a step (host launch overhead).   usage: python tools/gpu_train_profile.py [S|full] [batch]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.ops.loss import CharbonnierLoss  # noqa: E402
from fcvsr_b200.ops.optim import Adam  # noqa: E402
from fcvsr_b200.train import train_step  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "full"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
m = (arch.GShiftNet if variant == "full" else arch.GShiftNet_S)().to(dev).train()
m.load_state_dict(arch.seeded_state_dict(variant, 0))
opt = Adam(m.parameters(), lr=5e-6, weight_decay=1e-5)
g = torch.Generator().manual_seed(1)
x = torch.rand(batch, 7, 1, 64, 64, generator=g).to(dev)
hr = torch.rand(batch, 1, 256, 256, generator=g).to(dev)
for _ in range(2):
    train_step(m, opt, x, hr, CharbonnierLoss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    train_step(m, opt, x, hr, CharbonnierLoss)
e1.record()
torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / 3:.1f} ms (device events, {variant}, batch {batch})")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train_step(m, opt, x, hr, CharbonnierLoss)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
if not ev:
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"device-busy {tot / 1e3:.1f} ms in {sum(e.count for e in ev)} kernel launches")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:25]:
    print(f"{e.self_device_time_total / 1e3:9.2f} ms  n={e.count:5d}  {e.key[:110]}")
