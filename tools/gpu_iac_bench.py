"""Same-box timing of one IAC iteration (both directions) at B x 180 x 320 x 64: the materialised-taps path (the F.1 1x1
convolution for all six iterations / 6 + fcvsr_iac_step) against fcvsr_iac_step_tc (taps on chip).  CUDA events, 20 reps."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcvsr_b200 import _capi as C


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    H, W = 180, 320
    dev = torch.device("cuda:0")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    prev32 = [r(B, H, W, 64) for _ in range(2)]
    prev16 = [p.bfloat16() for p in prev32]
    xin = [r(B, H, W, 64) for _ in range(2)]
    offs = 2.0 * r(B, H, W, 24)
    kp = r(B, H, W, 64).bfloat16()
    w = (0.1 * r(192, 64)).bfloat16()
    bias = 0.1 * r(192)
    taps = r(B, H, W, 6 * 192).half()
    nxt = [torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16) for _ in range(2)]

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    for p16 in (0, 1):
        pv = prev16 if p16 else prev32
        old = lambda: C.call("fcvsr_iac_step", pv[0].data_ptr(), 64, pv[1].data_ptr(), 64, xin[0].data_ptr(), 64, xin[1].data_ptr(),
                             64, nxt[0].data_ptr(), 64, nxt[1].data_ptr(), 64, offs.data_ptr(), 24, 4, 6, taps.data_ptr(), 1152, 1,
                             B, H, W, 2 | (4 if p16 else 0), st)
        new = lambda: C.call("fcvsr_iac_step_tc", pv[0].data_ptr(), 64, pv[1].data_ptr(), 64, p16, xin[0].data_ptr(), 64,
                             xin[1].data_ptr(), 64, nxt[0].data_ptr(), 64, nxt[1].data_ptr(), 64, offs.data_ptr(), 24, 4, 6,
                             kp.data_ptr(), 64, w.data_ptr(), bias.data_ptr(), B, H, W, st)
        print(f"B={B} prev16={p16}: iac_step {t(old):.1f} us   iac_step_tc {t(new):.1f} us", flush=True)


if __name__ == "__main__":
    main()
