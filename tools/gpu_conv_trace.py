"""Bring-up helper: per-tile pipeline timeline of CTA 0 of conv_tc2 (FCVSR_TC_DBG=16)."""
import ctypes
import os
import sys

os.environ["FCVSR_TC_DBG"] = os.environ.get("FCVSR_TC_DBG", "16")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C  # noqa: E402
from fcvsr_b200.engine import _ConvPack  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
B, ci, co, H, W = 4, 64, 64, 180, 320
x = torch.randn(B, H, W, ci, device=dev)
w = torch.randn(co, ci, 3, 3, device=dev) / 24
pk = _ConvPack(w, None)
y = torch.empty(B, H, W, co, device=dev)
a2 = (x.data_ptr(), ci, pk.w_tc.data_ptr(), 9 * ci, 0, 0, 0, 0, 0, y.data_ptr(), co, B, H, W, ci, co, 0, 0.0, 0, 0, 0, 0, 0, 0, st)
for _ in range(3):
    C.call("fcvsr_conv3x3_tc_resident", *a2)
torch.cuda.synchronize()
n = 16 * 16
buf = (ctypes.c_longlong * n)()
lib = C.lib()
lib.fcvsr_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.fcvsr_debug_conv_trace(buf, n)
v = list(buf)
t0 = min(t for t in v if t > 0)
names = ["ld0_wait", "ld0_iss", "ld1_wait", "ld1_iss", "mma_tmE", "aF0", "aF1", "mma_done", "epi_tmF", "epi_done", "ldw0", "st0", "ldw3"]
print("clk since first stamp, CTA 0, per tile:", " ".join(f"{n:>9s}" for n in names))
for t in range(15):
    row = v[t * 16: t * 16 + 13]
    print(f"tile {t:2d}:                              ", " ".join(f"{(r - t0) if r else 0:9d}" for r in row))
