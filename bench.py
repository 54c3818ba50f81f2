#!/usr/bin/env python
"""FCVSR x4 super-resolution throughput benchmark (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--variant full|S]

Workload at N=1 (BASELINE.json configs[1]): FCVSR (GShiftNet) forward on synthetic 7-frame 180x320 clips,
x4 -> 720x1280, random-init weights (fcvsr_b200.arch.seeded_state_dict).  A step = one forward over a
batch of B independent 7-frame windows = B output frames.  With N > 1 (torchrun, one rank per GPU) every
rank processes its own B windows per step (windows are independent: no data-path collective, weak scaling).

Prints ONE JSON line (rank 0):
  value / e2e     headline mode (--dtype, default bf16 operands): frames/s with the clip resident in HBM (CUDA events, max over
                  ranks) and through the public API from pinned host memory (H2D + forward + D2H of every step inside the timed region,
                  transfers on side streams so that they overlap the neighbouring steps' forwards)
  modes           the same two numbers for BOTH arithmetic modes of BASELINE config 2: "tf32" (fp32 storage, TF32 tensor-core
                  operands: the contract's fp32 mode, max-abs <= 1e-3) and "bf16"
  roofline        dominant kernel (tcgen05 3x3 implicit-GEMM conv) timed INSIDE THE GRAPH-REPLAYED STEP: event-record nodes around
                  every launch of an instrumented capture of the same launch sequence (Engine.profile_graph_replay)
  sequence        BASELINE config 3: a 100-frame sequence sharded over the ranks by output-frame range with LR halos (strong scaling)
  train           BASELINE config 4: fwd + bwd + Adam on per-GPU batch 8 of 7x64x64 crops, NCCL gradient all-reduce timed separately
  cpu_baseline    the oracle port of the reference forward on the host cores (N = 1 only)
`--impl reference` times that CPU arm alone under the same contract.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "x4 SR output frames/sec (180x320->720x1280)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=6, help="7-frame windows per step and per GPU (sweep on B200: 2 -> 210, 4 -> 234, 6 -> 242, 8 -> 239 frames/s)")
    ap.add_argument("--variant", default="full", choices=["full", "S"])
    ap.add_argument("--height", type=int, default=180)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--dtype", default="bf16", choices=["tf32", "bf16"],
                    help="headline tensor-core operand type: bf16 (fp32 accumulate and residual streams) or tf32 (fp32 storage; the "
                         "contract's fp32 mode); the other mode is measured too and reported under `modes`")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sequence", action="store_true", help="skip the config-3 sequence pass")
    ap.add_argument("--no-train", action="store_true", help="skip the config-4 training-step pass")
    ap.add_argument("--single-mode", action="store_true", help="measure only --dtype (profiling runs)")
    ap.add_argument("--seq-frames", type=int, default=100)
    ap.add_argument("--train-steps", type=int, default=10)
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def committed_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the first `conv_tc_kernel<3,...>` launch in the newest committed
    `ncu --set full` summary (profiles/r*_conv_tc_full_summary.txt): parsed, not a constant."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_tc_full_summary.txt")),
                   key=lambda f: int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cur, rd, wr = None, None, None
    for line in open(files[-1]):
        line = line.strip()
        if line.startswith("--- launch"):
            if cur and rd is not None and wr is not None:
                break
            cur, rd, wr = None, None, None
        elif line.startswith("Kernel Name:") and "conv_tc_kernel<3" in line:
            cur = line
        elif cur and line.startswith("dram__bytes_read.sum:"):
            v, u = line.split(":")[1].split()
            rd = float(v) * unit[u]
        elif cur and line.startswith("dram__bytes_write.sum:"):
            v, u = line.split(":")[1].split()
            wr = float(v) * unit[u]
    if cur and rd is not None and wr is not None:
        return int(rd + wr), os.path.relpath(files[-1], ROOT)
    return None, os.path.relpath(files[-1], ROOT)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        # one long-running nvidia-smi sampling every 50 ms (a fresh process per sample takes ~150 ms and would see a 0.4 s
        # timed region once or twice)
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                     "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        try:
            for line in proc.stdout:
                if self.stop_flag:
                    break
                line = line.strip()
                if line:
                    self.samples.append([t.strip() for t in line.split(",")])
        finally:
            proc.kill()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def cpu_arm(args, steps, warmup):
    """The reference forward's CPU port (oracle) on all host cores: seconds per step (one clip per step)."""
    from fcvsr_b200 import arch
    from oracle import fcvsr_oracle as O
    from oracle.make_golden import make_clip
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = arch.seeded_state_dict(args.variant, 0)
    x = make_clip(1234, 1, args.height, args.width)
    with torch.no_grad():
        for _ in range(warmup):
            O.forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(sd, x)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, cores


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    vname = "FCVSR" if args.variant == "full" else "FCVSR-S"
    base_workload = f"{vname} forward, synthetic 7-frame {args.height}x{args.width} clips x4"
    mode_text = {"tf32": "fp32 storage, TF32 tensor-core operands, fp32 accumulate",
                 "bf16": "bf16 operand tensors, fp32 accumulate and residual streams"}
    workload = base_workload + ", " + mode_text[args.dtype]

    if args.impl == "reference":
        if rank != 0:
            return
        dt, cores = cpu_arm(args, args.steps, args.warmup)
        fps = 1.0 / dt
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": base_workload + ", fp32 on the host CPU cores (oracle port of the reference forward)",
                           "batch_per_step": 1},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.steps} forwards of one 7x{args.height}x{args.width} clip"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from fcvsr_b200 import arch
    from oracle.make_golden import make_clip
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W = args.batch, args.height, args.width
    sd = arch.seeded_state_dict(args.variant, 0)
    cls = arch.GShiftNet if args.variant == "full" else arch.GShiftNet_S
    x_host = make_clip(1234 + rank, B, H, W).pin_memory()
    y_host = torch.empty(B, 1, 4 * H, 4 * W).pin_memory()
    x_dev = x_host.to(dev)
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()                   # side streams joined: every transfer of the timed steps ends before e1
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    models = {}

    def measure(dtype, sample_clocks):
        """frames/s of one arithmetic mode: resident (`value`) and end to end through model(x) from pinned host memory."""
        model = cls().to(dev).eval()
        model.load_state_dict(sd)
        model.compute_dtype = dtype
        models[dtype] = model
        with torch.no_grad():
            model(x_dev)                                   # builds packs / workspace, eager warm-up
            eng = model._engine
            eng.use_graph = not args.no_graph
            for _ in range(warm):
                model(x_dev)
            sampler = None
            if sample_clocks:
                sampler = ClockSampler(local)
                sampler.start()
            ms = timed(lambda: model(x_dev), args.steps)
            if sampler is not None:
                sampler.stop_flag = True
            # end to end: the H2D copy lands in the graph's own input buffer and the D2H copy reads its output buffer, so the
            # public call adds no device-side copies; both transfers and the forward are inside the timed region every step
            eng.clone_output = False
            xin = eng.static_input(B, H, W, dev) if eng.use_graph else None
            # A streaming caller overlaps the transfers of neighbouring steps with the forward: step i's clip travels on an
            # input stream into one of two device staging buffers while step i - 1 computes, and its result leaves through one
            # of two staging buffers on an output stream while step i + 1 computes (device-to-device moves of 10 + 22 MB between
            # the staging buffers and the graph's own input / output buffers).  Every H2D and D2H of the timed steps is inside
            # the timed region: the side streams are joined before the closing event.
            main = torch.cuda.current_stream()
            s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            x_stage = [torch.empty_like(x_dev) for _ in range(2)]
            y_stage = [torch.empty(B, x_dev.shape[2], 4 * H, 4 * W, device=dev) for _ in range(2)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_in_free = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]
            ev_out_free = [torch.cuda.Event() for _ in range(2)]
            counter = [0]

            def step_e2e():
                k = counter[0] & 1
                counter[0] += 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_in_free[k])                       # the staging buffer's previous content was consumed
                    x_stage[k].copy_(x_host, non_blocking=True)
                    ev_in[k].record(s_in)
                main.wait_event(ev_in[k])
                if xin is not None:
                    xin.copy_(x_stage[k], non_blocking=True)
                    ev_in_free[k].record(main)
                    y = model(xin)
                else:
                    y = model(x_stage[k])
                    ev_in_free[k].record(main)
                main.wait_event(ev_out_free[k])                          # the output staging buffer has left the device
                y_stage[k].copy_(y, non_blocking=True)
                ev_out[k].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_out[k])
                    y_host.copy_(y_stage[k], non_blocking=True)
                    ev_out_free[k].record(s_out)

            def join():
                main.wait_stream(s_in)
                main.wait_stream(s_out)

            for _ in range(2):
                step_e2e()
            join()
            ms_e2e = timed(step_e2e, args.steps, finish=join)
            eng.clone_output = True
        frames = B * world * args.steps
        return {"value": frames / (ms * 1e-3), "ms_per_step": ms / args.steps, "e2e": frames / (ms_e2e * 1e-3),
                "launches_per_step": eng.launches, "tc_launches_per_step": eng.tc_launches,
                "clocks": sampler.summary() if sampler is not None else None}

    order = [args.dtype] if args.single_mode else [args.dtype] + [d for d in ("tf32", "bf16") if d != args.dtype]
    res = {d: measure(d, d == args.dtype) for d in order}
    head = res[args.dtype]
    fps = head["value"]

    # ---- dominant-kernel roofline, measured inside the graph-replayed step -----------------------------------------------------
    roof = None
    if rank == 0:
        hbm, tfl, src = peaks()
        eng = models[args.dtype]._engine
        how = "event-record nodes around every launch of the graph-replayed step (instrumented capture of the same launch sequence)"
        try:
            if args.no_graph:
                raise RuntimeError("--no-graph")
            with torch.no_grad():
                prof, replay_ms = eng.profile_graph_replay(x_dev, reps=5)
        except Exception as e:                         # external events unavailable: eager per-launch events, said so
            how = f"eager per-launch CUDA events (graph instrumentation failed: {type(e).__name__})"
            with torch.no_grad():
                eng.use_graph = False
                models[args.dtype](x_dev)
                eng.profile = []
                models[args.dtype](x_dev)
                torch.cuda.synchronize()
                p0, eng.profile = eng.profile, None
                eng.use_graph = not args.no_graph
            prof = [(k, f, by, a.elapsed_time(b)) for (k, f, by, a, b) in p0]
            replay_ms = sum(t for _, _, _, t in prof)
        # the dominant kernel is the 3x3 instantiation conv_tc_kernel<3, .>; the 1x1 instantiation <1, .> is a separate,
        # memory-bound kernel and is not folded into this tensor roofline
        tc = [(f, t) for (k, f, _, t) in prof if k.startswith("tc") and " k3 " in k]
        t_tc = sum(t for _, t in tc)
        fl_tc = sum(f for f, _ in tc)
        ach = fl_tc / (t_tc * 1e-3) / 1e12 if t_tc > 0 else 0.0
        traffic, traffic_src = committed_traffic()
        by_kernel = {}
        for (k, f, _, t) in prof:
            kk = k.split(" ")[0] if k.startswith("tc") else k
            if k.startswith("tc"):
                kk = "conv_tc k3" if " k3 " in k else "conv_tc k1"
            a = by_kernel.setdefault(kk, [0, 0.0])
            a[0] += 1
            a[1] += t
        top = sorted(by_kernel.items(), key=lambda kv: -kv[1][1])[:8]
        roof = {"kernel": f"conv_tc_kernel<3> (tcgen05 {args.dtype} implicit-GEMM 3x3 convolution, every launch of the step)",
                "bound": "tensor", "achieved": ach, "peak": tfl, "unit": "TFLOP/s", "frac": ach / tfl if tfl else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": src, "how": how,
                "launches": len(tc), "avg_launch_us": 1e3 * t_tc / max(len(tc), 1),
                "instrumented_replay_ms": replay_ms, "production_replay_ms": head["ms_per_step"],
                # sum of this kernel's in-graph durations over the step time: the pyramid levels run concurrently, so the sum
                # of all kernels' durations exceeds the step time
                "kernel_time_over_step_time": t_tc / replay_ms if replay_ms else None,
                "top_kernels_ms": {k: {"n": n, "ms": round(t, 4)} for k, (n, t) in top},
                "note": ("TF32 operands run at half the bf16 tensor rate; " if args.dtype == "tf32" else "") +
                        "peak is the measured sustained bf16 dense figure"}
        # whole-forward roofline of SURVEY 8(d): live conv FLOPs per LR pixel (22.50 M FCVSR / 9.61 M FCVSR-S) on the tensor
        # pipe plus the compulsory HBM bytes of the non-GEMM stages (0.48 GB per 180x320 FCVSR frame) -- the
        # frames/s the forward could reach if every kernel sat on its own roofline; per GPU
        mflop_px = 22.50 if args.variant == "full" else 9.61
        t_tensor = mflop_px * 1e6 * H * W / (tfl * 1e12)
        # bytes per LR pixel: 3 x MGAAbk (768 in + 16A offsets out; IAC 768 + 16A in, 512 out) + MFFR 512 + up-sampler 1092
        bytes_px = 7748 + 96 * (6 if args.variant == "full" else 3)
        t_hbm = bytes_px * H * W / (hbm * 1e9)
        roof["forward"] = {"alg_tflop_per_frame": mflop_px * 1e6 * H * W / 1e12, "alg_gb_per_frame": bytes_px * H * W / 1e9,
                           "bound_frames_per_s": 1.0 / (t_tensor + t_hbm), "frac": (fps / world) * (t_tensor + t_hbm)}

    # ---- BASELINE config 3: fixed 100-frame sequence, output-frame ranges sharded over the ranks (strong scaling) --------------
    seq_res = None
    if not args.no_sequence and args.variant == "full":
        from fcvsr_b200.sequence import shard_range, super_resolve_sequence
        n = args.seq_frames
        g = torch.Generator().manual_seed(4321)
        seq = (torch.round(255 * torch.rand(n, 1, H, W, generator=g)) / 255).pin_memory()
        lo, hi = shard_range(n, rank, world)
        out_host = torch.empty(hi - lo, 1, 4 * H, 4 * W).pin_memory()
        model = models[args.dtype]

        def run_seq():
            super_resolve_sequence(model, seq, batch=B, rank=rank, world=world, stream_chunk=4 * B, out=out_host)

        with torch.no_grad():
            run_seq()
            ms_seq = min(timed(run_seq, 1) for _ in range(2))
        seq_res = {"value": n / (ms_seq * 1e-3), "unit": "frames/s", "frames": n, "scaling": "strong", "ms_per_sequence": ms_seq,
                   "dtype": args.dtype, "windows_per_launch": B,
                   "what": f"FCVSR over one synthetic {n}-frame {H}x{W} sequence: output-frame ranges sharded over {world} rank(s) "
                           "with LR halos, replicate edge padding, pinned-memory streaming (H2D of LR chunks and D2H of HR frames "
                           "inside the timed region), device time, max over ranks"}

    # ---- BASELINE config 4: training step with the NCCL gradient all-reduce -----------------------------------------------------
    train_res = None
    if not args.no_train:
        try:
            from fcvsr_b200.train import bench_train_step
            train_res = bench_train_step(dev, rank, world, steps=args.train_steps, variant=args.variant)
        except Exception as e:       # the inference headline must not depend on the training arm
            train_res = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    nbytes_in, nbytes_out = x_host.numel() * 4, y_host.numel() * 4
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "batch_per_step_per_gpu": B, "parallelism": f"window-sharded x{world}",
                       "l2": "working set per step (~1.5 GB of NHWC feature maps) exceeds the 126 MB L2; no flush needed",
                       "cuda_graph": not args.no_graph},
            "clocks": head["clocks"],
            "e2e": {"value": head["e2e"], "unit": "frames/s", "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": nbytes_out,
                    "how": "every step: H2D of its clip from pinned host memory, model(x), D2H of its result; the transfers run on an "
                           "input and an output stream through two device staging buffers each, so those of neighbouring steps overlap "
                           "the forward; the streams are joined before the closing event"},
            "modes": {d: {"value": r["value"], "e2e": r["e2e"], "ms_per_step": r["ms_per_step"], "unit": "frames/s",
                          "arithmetic": mode_text[d],
                          "tolerance": "max-abs <= 1e-3, |dPSNR| <= 0.01 dB" if d == "tf32" else "max-abs <= 5e-3, PSNR >= 60 dB"}
                      for d, r in res.items()},
            "gpu_launches": head["launches_per_step"] * args.steps,
            "roofline": roof, "sequence": seq_res, "train": train_res}
    if not args.no_cpu_baseline and world == 1:
        dt, cores = cpu_arm(args, 2, 1)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"2 timed forwards (after 1 warm-up) of one 7x{H}x{W} clip, oracle port on CPU"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
