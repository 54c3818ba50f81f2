"""Time fcvsr_conv3x3_c64_to1 (conv_last0 + skip, bf16 input) at B x 720 x 1280.  usage: gpu_last_bench.py [B]"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = 720, 1280
dev = torch.device("cuda:0")
x = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
res = torch.randn(B, H, W, device=dev)
y = torch.empty(B, H, W, device=dev)
wh = (ctypes.c_float * 576)(*[0.01 * i for i in range(576)])
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
f = lambda: C.call("fcvsr_conv3x3_c64_to1", x.data_ptr(), 64, ctypes.addressof(wh), 0.25, res.data_ptr(), y.data_ptr(), B, H, W, st)
for _ in range(3): f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"B{B}: {us:.1f} us  {B*H*W*128/us/1e3:.0f} GB/s of input")
