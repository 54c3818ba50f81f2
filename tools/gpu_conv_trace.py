"""Bring-up helper: per-tile pipeline timeline of CTA 0 of conv_tc (v1) or conv_tc2 (v2), built separately with
-DTC_TRACE / -DT2_TRACE.   usage: [KS=1|3] [HW=180x320] python tools/gpu_conv_trace.py [v1|v2] [bf16|tf32] [Cin] [Cout] [B]"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcvsr_b200.engine import _ConvPack  # noqa: E402

ver = sys.argv[1] if len(sys.argv) > 1 else "v1"
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
ci = int(sys.argv[3]) if len(sys.argv) > 3 else 64
co = int(sys.argv[4]) if len(sys.argv) > 4 else 64
B = int(sys.argv[5]) if len(sys.argv) > 5 else 4
if ver != "v1":
    sys.exit("conv_tc2 (v2) was removed in round 2; only v1 = conv_tc.cu remains")
src = "conv_tc.cu"
so = os.path.join(ROOT, "gpurun_out", f"libtrace_{ver}.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
extra = [f"-D{d}" for d in os.environ.get("T2_DEFS", "").split()]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DTC_TRACE", "-DFCVSR_BRINGUP", *extra,
                       "-shared", "-Xcompiler", "-fPIC", "--cudart", "shared", os.path.join(ROOT, "fcvsr_b200/csrc", src), "-o", so,
                       "-Wno-deprecated-gpu-targets"])
lib = ctypes.CDLL(so)
op16 = mode == "bf16"
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
H, W = [int(v) for v in os.environ.get("HW", "180x320").split("x")]
x = torch.randn(B, H, W, ci, device=dev)
KS = int(os.environ.get("KS", "3"))
w = torch.randn(co, ci, KS, KS, device=dev) / (KS * KS * ci) ** 0.5
bias = torch.randn(co, device=dev)
pk = _ConvPack(w, None, op16=op16)
if op16:
    x = x.to(torch.bfloat16)
f32out = bool(os.environ.get("F32OUT"))
y = torch.empty(B, H, W, co, device=dev, dtype=torch.bfloat16 if (op16 and not f32out) else torch.float32)
y2 = torch.empty(B, H, W, co, device=dev, dtype=torch.bfloat16 if op16 else torch.float32) if f32out else None
V, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
if ver == "v1":
    fn = lib.fcvsr_conv2d_tc
    fn.argtypes = [V, I, V, V, V, I, V, I, V, I, I, I, I, I, I, I, I, F, V, I, V, I, I, I, I, V]
    args = (x.data_ptr(), ci, pk.w_tc.data_ptr(), bias.data_ptr(), None, 0, None, 0, y.data_ptr(), co, B, H, W, ci, co, KS, 2, 0.1,
            None, 0, y2.data_ptr() if f32out else None, co if f32out else 0, 0 if f32out else (1 if op16 else 0), 0, int(op16), st)
else:
    fn = lib.fcvsr_conv3x3_tc_resident
    fn.argtypes = [V, I, V, I, V, V, I, V, I, V, I, I, I, I, I, I, I, F, V, I, V, I, I, I, I, V]
    args = (x.data_ptr(), ci, pk.w_tc.data_ptr(), 9 * ci, bias.data_ptr(), None, 0, None, 0, y.data_ptr(), co, B, H, W, ci, co, 2, 0.1,
            None, 0, None, 0, 1 if op16 else 0, 0, int(op16), st)
for _ in range(3):
    rc = fn(*args)
    assert rc == 0, rc
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    fn(*args)
e1.record()
torch.cuda.synchronize()
print(os.environ.get("T2_DEFS", ""), f"{ver} {mode} {ci}->{co} B{B}: {e0.elapsed_time(e1) * 5:.1f} us per launch (traced build)")
n = 64 * 16
buf = (ctypes.c_longlong * n)()
lib.fcvsr_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.fcvsr_debug_conv_trace(buf, n) == 0
v = list(buf)
clk = v[63 * 16 + 14] - v[62 * 16 + 14]
ns = v[63 * 16 + 15] - v[62 * 16 + 15]
print(f"CTA 0 body: {clk} clk in {ns} ns -> clock64 rate {clk / max(ns, 1) * 1e3:.0f} MHz")
for t_ in (62, 63):
    for s_ in range(16):
        v[t_ * 16 + s_] = 0
t0 = min(t for t in v if t > 0)
names = {0: "pr_aE", 1: "pr_iss", 4: "mma_tmE", 5: "aF_first", 6: "aF_last", 7: "mma_iss", 8: "epi_tmF", 10: "epi_bar", 11: "epi_ld", 12: "epi_math", 9: "epi_done"}
print("clk since first stamp, CTA 0, per tile: ", " ".join(f"{names[k]:>9s}" for k in names))
for t in range(16):
    row = [v[t * 16 + k] for k in names]
    if not any(row):
        break
    print(f"tile {t:2d}:                                ", " ".join(f"{(r - t0) if r else 0:9d}" for r in row))
