import os, sys, torch
sys.path.insert(0, os.getcwd())
from fcvsr_b200.ops.dcn import modulated_deform_conv
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
B, H, W = 1, 180, 320
x = torch.randn(B, 64, H, W, generator=g).to(dev)
w = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
b = torch.randn(64, generator=g).to(dev)
sig = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
off = (sig * torch.randn(B, 288, H, W, generator=g)).to(dev)
msk = torch.rand(B, 144, H, W, generator=g).to(dev)
with torch.no_grad():
    for _ in range(5):
        modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16)
    e1.record(); torch.cuda.synchronize()
print(f"sigma {sig}: {e0.elapsed_time(e1)*100:.1f} us per call")
