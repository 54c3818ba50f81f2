"""One eager pass over every kernel that profiles/r1_conv_tc_full / r1_fft_iac_full do not cover, for an ncu capture:

    ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:'<names below>' \
        -o gpurun_out/r1_misc python tools/ncu_misc_step.py

FCVSR at 180x320, 4 windows, bf16 mode, with SCGroupN = 1 (same kernels and shapes as the full model, a tenth of the SCNet
launches), then the DCN forward / backward entries and the Charbonnier loss at the SURVEY 8(a10) / config-4 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
import fcvsr_b200.ops.dcn as dcn  # noqa: E402
from fcvsr_b200.ops.loss import CharbonnierLoss  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dev = torch.device("cuda:0")
sd = arch.seeded_state_dict("full", 0, SCGroupN=1)
m = arch.GShiftNet(SCGroupN=1).to(dev).eval()
m.load_state_dict(sd)
m.compute_dtype = "bf16"
x = make_clip(1234, 4, 180, 320).to(dev)
with torch.no_grad():
    m(x)
    m._engine.use_graph = False
    m._engine.multi_stream = False
    torch.cuda.synchronize()
    torch.cuda.profiler.start()                 # ncu --profile-from-start off: capture from here
    y = m(x)
torch.cuda.synchronize()

g = torch.Generator().manual_seed(0)
B, H, W = 1, 180, 320
xi = torch.randn(B, 64, H, W, generator=g).to(dev).requires_grad_()
w = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev).requires_grad_()
b = torch.randn(64, generator=g).to(dev).requires_grad_()
off = (2.0 * torch.randn(B, 288, H, W, generator=g)).to(dev).requires_grad_()
msk = torch.rand(B, 144, H, W, generator=g).to(dev).requires_grad_()
for prec in ("tf32", "fp32"):
    dcn.PRECISION = prec
    for fast in (True, False):
        dcn.BACKWARD_NHWC = fast
        out = dcn.modulated_deform_conv(xi, off, msk, w, b, 1, 1, 1, 1, 16)
        if prec == "fp32":
            out.square().sum().backward()
torch.cuda.synchronize()
sr = torch.rand(8, 1, 256, 256, device=dev, requires_grad=True)
hr = torch.rand(8, 1, 256, 256, device=dev)
CharbonnierLoss(sr, hr).backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(y.abs().mean()))
