"""Four eager config-4 training steps (no profiler, no graph): the target of the ncu captures of the backward kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.ops.loss import CharbonnierLoss  # noqa: E402
from fcvsr_b200.ops.optim import Adam  # noqa: E402
from fcvsr_b200.train import train_step  # noqa: E402

dev = torch.device("cuda:0")
m = arch.GShiftNet().to(dev).train()
m.load_state_dict(arch.seeded_state_dict("full", 0))
opt = Adam(m.parameters(), lr=5e-6, weight_decay=1e-5)
g = torch.Generator().manual_seed(1)
x = torch.rand(8, 7, 1, 64, 64, generator=g).to(dev)
hr = torch.rand(8, 1, 256, 256, generator=g).to(dev)
for _ in range(4):
    loss = train_step(m, opt, x, hr, CharbonnierLoss)
torch.cuda.synchronize()
print("loss", float(loss))
