#!/usr/bin/env python
"""bench.py -- images/sec of the ResNet-34 160x384 waypoint train step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation of the same step

Workload (config.workload = "config2"): BASELINE.json configs[1] -- ImagePolicyModelSS('resnet34'), synthetic
batch 256/GPU, one full training step = student forward (train-mode BN) + phase-0 L1 waypoint loss against fixed
image-space targets + backward + Adam(lr=1e-4); data-parallel replicas all-reduce the gradient once per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GFLOP_PER_TRAIN_IMG = 28.034      # 3 x 9.441116 - 0.289014 (no dgrad for the stem)
METRIC = "images/sec (ResNet34 160x384 waypoint train step)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampled every 50 ms in a side process.  It is started before the warm-up (nvidia-smi needs a few hundred
    ms to come up) and only the samples stamped inside [mark(), stop()] -- the timed regions -- are reported."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):             # a timed region starts
        self.windows.append([time.time(), None])

    def end(self):              # ... and ends
        if self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.time()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.end()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], None, [], set()
        for stamp, ln in self.lines:
            # a line is printed right after its sample is taken: keep those received inside the marked window
            if self.windows and not any(a <= stamp <= b + 0.02 for a, b in self.windows):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_min_mhz=(sm[0] if sm else None), sm_max_mhz=mx,
                    power_w_max=(max(power) if power else None), reasons=sorted(reasons), samples=len(sm),
                    window="device-resident timed loop + the two end-to-end timed loops")


def synthetic(B, seed, torch):
    g = torch.Generator().manual_seed(seed)
    rgb_u8 = torch.randint(0, 256, (B, 3, 160, 384), dtype=torch.uint8, generator=g)
    rgb = rgb_u8.float() / 255
    speed = torch.rand(B, generator=g) * 10
    cmd = torch.randint(1, 5, (B,), generator=g).float()
    target = torch.rand(B, 5, 2, generator=g) * torch.tensor([384.0, 160.0])
    synthetic.last_u8 = rgb_u8
    return rgb, speed, cmd, target


def cpu_threads():
    """Threads for the CPU arm.  Measured on the 128-thread GPU host: torch/oneDNN on this B=8 step is ~20x SLOWER with
    128 intra-op threads (0.6 img/s) than with a few dozen, so the arm uses min(cores, LBC_CPU_THREADS or 32)."""
    n = os.cpu_count() or 1
    return max(1, min(n, int(os.environ.get("LBC_CPU_THREADS", "32"))))


def cpu_step_rate(torch, B, warm, iters, threads=None):
    """The reference's CPU implementation of the step (oracle port: torch CPU kernels + Adam), images/s."""
    import lbc_oracle as orc
    import learningbycheating_b200 as lbc
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = orc.leafify(lbc.ImagePolicyModelSS("resnet34", all_branch=True).state_dict())
    rgb, speed, cmd, target = synthetic(B, 1, torch)
    oh = orc.one_hot(cmd)
    st = orc.new_adam_state()

    def step():
        for k in orc.param_keys(sd):
            sd[k].grad = None
        pred, _, newbuf = orc.policy_forward(sd, rgb, speed, oh, "resnet34", True, True)
        loss = orc.phase0_loss(pred, target).mean()
        loss.backward()
        st["step"] += 1
        with torch.no_grad():
            pk = [k for k in orc.param_keys(sd) if sd[k].grad is not None]
            for k in pk:
                if k not in st["m"]:
                    st["m"][k] = torch.zeros_like(sd[k])
                    st["v"][k] = torch.zeros_like(sd[k])
            orc.adam_step({k: sd[k] for k in pk}, {k: sd[k].grad for k in pk}, st["m"], st["v"], st["step"])
            for k, v in newbuf.items():
                sd[k] = v
        return float(loss.detach())

    for _ in range(warm):
        step()
    t0 = time.time()
    for _ in range(iters):
        step()
    dt = time.time() - t0
    return B * iters / dt, dt / iters


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    torch.set_num_threads(cores)
    B = 8
    rate, sec = cpu_step_rate(torch, B, max(1, min(args.warmup, 2)), max(1, args.steps), cores)
    sample = ("oracle port (torch CPU fp32) of the config2 step on a bounded sample: batch %d per step, %d of %d host "
              "threads (more threads are slower for this batch)" % (B, cores, os.cpu_count() or 1))
    out = dict(impl="reference", metric=METRIC, value=rate, unit="images/s", n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32", data="synthetic",
               config=dict(workload="config2", step="student fwd+bwd+Adam, phase-0 L1 vs fixed targets", batch_per_step=B),
               cpu_baseline=dict(value=rate, unit="images/s", cores=torch.get_num_threads(), kind="port", sample=sample),
               e2e=dict(value=rate, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib, train_image_phase0 as p0
    from learningbycheating_b200.distributed import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200; there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    pair = args.pair if args.pair >= 0 else int(os.environ.get("LBC_PAIR", "7") or 0)   # bit 0: CTA-pair GEMMs, bit 1: wgrad3
    L.lbc_set_fast_kernels((0 if args.no_fast else 1) | (4 if pair & 1 else 8) | (16 if pair & 2 else 32) | (64 if pair & 4 else 128))
    B = args.batch
    torch.manual_seed(0)
    net = lbc.ImagePolicyModelSS("resnet34", all_branch=True, lbc_precision=args.precision).to(dev).train()
    opt = lbc.Adam(net.parameters(), lr=1e-4)
    dp = DataParallel(net, opt)
    crit = p0.LocationLoss(device=dev)
    rgb_h, speed_h, cmd_h, target_h = synthetic(B, 1 + rank, torch)
    rgb, speed, target = rgb_h.to(dev), speed_h.to(dev), target_h.to(dev)
    oh = lbc.one_hot(cmd_h).to(dev)

    def step():
        pred, _ = net(rgb, speed, oh)
        loss = crit(pred, target).mean()
        opt.zero_grad()
        loss.backward()
        dp.step_after_backward()
        return loss

    step()                      # builds the native engine (allocations) outside every timed region
    dp.sync_initial_state()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()         # comes up during the warm-up; only samples after sampler.mark() are reported
    for _ in range(args.warmup):
        step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    if rank == 0:
        sampler.mark()
    n0 = L.lbc_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    if rank == 0:
        sampler.end()
    ms = e0.elapsed_time(e1)
    launches = L.lbc_kernel_launch_count() - n0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers (pinned): every step's inputs are copied host->device
    # inside the timed region (on the package's CudaPrefetcher side stream, overlapping the previous step's compute)
    # and the step's loss is read back device->host.
    from learningbycheating_b200.data import CudaPrefetcher
    # frames as the data collector stores them: uint8 [B,160,384,3] (data_collector.py:234-252); ToTensor's /255 runs on
    # the device (lbc_net_forward_u8), so the per-step H2D copy is 47 MB instead of 189 MB.  The fp32-frame variant (what
    # the reference's DataLoader hands over) is measured too and reported as e2e.fp32_frames.
    rgb_u8_p = synthetic.last_u8.permute(0, 2, 3, 1).contiguous().pin_memory()
    rgb_p, speed_p, target_p = rgb_h.pin_memory(), speed_h.pin_memory(), target_h.pin_memory()
    e2e_steps = max(5, min(args.steps, 40))    # same K as the device-resident loop (pipeline fill/drain is inside the region)

    oh_p = lbc.one_hot(cmd_h).pin_memory()     # pinned once, as a DataLoader(pin_memory=True) thread would hand it over

    def host_batches(n, frames):
        for _ in range(n):
            yield (frames, speed_p, oh_p, target_p)

    def e2e_run(n, frames):
        # every step's loss is read back to the host; the read of step i is issued after step i+1 has been enqueued
        # (as a logging loop would do), so the device never drains waiting for the host
        last, pending = None, None
        for r, s_, c, tg in CudaPrefetcher(host_batches(n, frames), dev):
            pred, _ = net(r, s_, c)
            l = crit(pred, tg).mean()
            opt.zero_grad()
            l.backward()
            dp.step_after_backward()
            if pending is not None:
                last = pending.item()  # device -> host read of the previous step's loss
            pending = l.detach()
        if pending is not None:
            last = pending.item()
        return last

    def e2e_measure(frames):
        e2e_run(3, frames)
        barrier()
        if rank == 0:
            sampler.mark()
        e0.record()
        e2e_run(e2e_steps, frames)
        e1.record()
        barrier()
        if rank == 0:
            sampler.end()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / (float(tt.item()) / 1e3)

    e2e_run(2, rgb_p)               # both variants warmed (pinned-host / side-stream allocator pools) before either is timed
    e2e_run(2, rgb_u8_p)
    e2e_fp32 = e2e_measure(rgb_p)
    e2e_value = e2e_measure(rgb_u8_p)
    clocks = sampler.stop() if rank == 0 else None
    small = speed_p.numel() * 4 + B * 4 * 4 + target_p.numel() * 4
    h2d = rgb_u8_p.numel() + small
    h2d_fp32 = rgb_p.numel() * 4 + small

    # ---- per-category device timing (CUDA events on the launching stream) for the roofline of the dominant kernel
    roof = None
    cats = {}
    nprof = 2
    if rank == 0:
        L.lbc_prof_reset()
        L.lbc_prof_enable(1)
    for _ in range(nprof):      # every rank steps (the step contains the gradient all-reduce); rank 0 records
        step()
    torch.cuda.synchronize()
    if rank == 0:
        import ctypes
        peaks = load_peaks()
        L.lbc_prof_enable(0)
        for cat in ("conv_fwd", "conv_dgrad", "conv_wgrad", "bn_fwd", "bn_bwd"):
            msd, cnt, fl, by = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
            _lib.check(L.lbc_prof_get(cat.encode(), ctypes.byref(msd), ctypes.byref(cnt), ctypes.byref(fl), ctypes.byref(by)))
            cats[cat] = dict(ms_per_step=msd.value / nprof, launches_per_step=cnt.value / nprof,
                             gflop_per_step=fl.value / nprof / 1e9, gbytes_per_step=by.value / nprof / 1e9)
        L.lbc_prof_reset()
        conv_ms = sum(cats[c]["ms_per_step"] for c in ("conv_fwd", "conv_dgrad", "conv_wgrad"))
        conv_gf = sum(cats[c]["gflop_per_step"] for c in ("conv_fwd", "conv_dgrad", "conv_wgrad"))
        achieved = conv_gf / conv_ms if conv_ms > 0 else 0.0     # GFLOP/ms == TFLOP/s
        bn_ms = cats["bn_fwd"]["ms_per_step"] + cats["bn_bwd"]["ms_per_step"]
        bn_gb = cats["bn_fwd"]["gbytes_per_step"] + cats["bn_bwd"]["gbytes_per_step"]
        bn_gbs = bn_gb / bn_ms * 1e3 if bn_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_ncu_summary.json")
        if os.path.exists(tpath):       # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            try:
                dk = json.load(open(tpath)).get("dominant_kernel")
                traffic, traffic_src = dk["dram_total"], dict(kernel=dk["kernel"], capture=dk["capture"],
                                                             algorithmic_bytes=dk["algorithmic_bytes"])
            except Exception:
                traffic = None
        roof = dict(bound="tensor", kernel="tcgen05 implicit-GEMM convolutions (fwd + dgrad + wgrad, all layers; CUDA-event "
                                           "brackets on the launching stream)",
                    achieved=achieved, peak=peaks["tf_sust"], unit="TFLOP/s", frac=achieved / peaks["tf_sust"],
                    traffic=traffic, traffic_of=traffic_src, peak_source=peaks["source"] + " bf16_tflops_sustained (kernels timed inside a long step)",
                    hbm_kernels=dict(kernel="BatchNorm statistics/apply/backward kernels", achieved=bn_gbs, peak=peaks["hbm"],
                                     unit="GB/s", frac=bn_gbs / peaks["hbm"], bytes="algorithmic (DESIGN.md section 3)"),
                    per_category=cats)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads()
        rate, sec = cpu_step_rate(torch, 8, 1, 4, cores)
        cpu_base = dict(value=rate, unit="images/s", cores=torch.get_num_threads(), kind="port",
                        sample="oracle port (torch CPU fp32) of the same step, batch 8, 4 timed iterations, %d of %d host threads"
                               % (cores, os.cpu_count() or 1))

    if rank == 0:
        out = dict(metric=METRIC, value=value, unit="images/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype=("bf16" if args.precision == "bf16" else "f32"), data="synthetic",
                   config=dict(workload="config2", model="ImagePolicyModelSS resnet34 160x384", global_batch=world * B,
                               batch_per_gpu=B, parallelism="dp%d" % world,
                               step="student fwd+bwd+Adam, phase-0 L1 vs fixed targets",
                               l2="inputs+activations per step (>5 GB) far exceed the 126 MB L2; no explicit flush",
                               fast_kernels=not args.no_fast, cta_pair_gemm=bool(pair & 1), wgrad_row_of_taps=bool(pair & 2), wgrad_cta_pair=bool(pair & 4)),
                   e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, steps=e2e_steps,
                            frames="uint8 NHWC host frames (pinned), copied on the prefetch stream every step",
                            fp32_frames=dict(value=e2e_fp32, h2d_bytes_per_step=h2d_fp32)),
                   gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu_base,
                   step_tflops=value * GFLOP_PER_TRAIN_IMG / 1e3, last_loss=float(loss.detach()))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32tc"])
    ap.add_argument("--no-fast", action="store_true", help="correctness-first kernels only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pair", type=int, default=-1, help="kernel variants: bit 0 = CTA-pair (cta_group::2) implicit-GEMM kernels, "
                                                         "bit 1 = row-of-taps weight gradient (default: LBC_PAIR)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
