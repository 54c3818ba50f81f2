"""Forward / backward time of the odd-shaped convolutions of the training step (config 4: batch 8 of 7x64x64) through
fcvsr_b200.autograd.conv2d, one shape at a time (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import autograd as A
dev = torch.device("cuda:0")
shapes = [("ConvBlk 4->4 k11", 16, 4, 4, 64, 33, 11, 1), ("ConvBlk 4->4 k9", 16, 4, 4, 64, 33, 9, 1), ("ConvBlk 4->4 k5", 16, 4, 4, 64, 33, 5, 1),
          ("ConvBlk 4->4 k1", 16, 4, 4, 64, 33, 1, 1), ("conv_last0 64->1", 8, 64, 1, 256, 256, 3, 1), ("feat_extract 7->448", 8, 7, 448, 64, 64, 3, 1),
          ("head 64->4 k1", 16, 64, 4, 64, 33, 1, 1), ("fuse 84->64", 8, 84, 64, 64, 64, 3, 1), ("L2_2 80->64", 8, 80, 64, 32, 32, 3, 1),
          ("rconcat 64->64 s2", 8, 64, 64, 64, 64, 3, 2), ("trunk 64->64", 8, 64, 64, 64, 64, 3, 1)]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, B, ci, co, H, W, k, s in shapes:
    x = torch.randn(B, ci, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5).requires_grad_(True)
    b = None if name.startswith("ConvBlk") else torch.randn(co, device=dev).requires_grad_(True)      # ConvBlk: bias=False
    def fwd():
        with torch.no_grad():
            return A.conv2d(x, w, b, s, "tf32")
    def fb():
        y = A.conv2d(x, w, b, s, "tf32")
        y.backward(torch.ones_like(y))
    tf, tfb = t(fwd), t(fb)
    print(f"{name:22s} B{B} {H}x{W}: fwd {tf:8.1f} us   fwd+bwd {tfb:8.1f} us")
