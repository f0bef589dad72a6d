#!/usr/bin/env python
"""bench.py — headline benchmark of the gs-dynamics hot path on B200 (contract: see DESIGN.md §Measurement).

Workload (BASELINE.json configs[2], the size the north-star target names): steady-state tracking iteration of train_gs.py —
get_loss (fused RGB+seg render, L1+SSIM, rigid/rot/iso/floor/bg priors) + backward + Adam — at G = 100 000 Gaussians on the
4 demo cameras @640x480, one random camera per iteration.  A "step" is one such iteration.  metric = tracking iters/sec.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|upstream_structure] [--gaussians G]

ours               : CUDA-graph replay of the iteration (inputs resident in HBM) -> value;  e2e: the same iteration through the
                     public API with the step's camera image + seg mask (8-bit, as the dataset's PNGs hold them) copied from pinned host
                     memory, unpacked on the device, and the loss read back, all inside
                     the timed region.  Side objects (N = 1): roofline (dominant kernel vs HBM and issue-slot peaks, the §8(d)
                     aggregate, per-stage table), episode (configs[2]: frames of 2 000 iterations with per-frame target upload and
                     initialize_per_timestep), secondary_50k (configs[1]), render_1280x720 (A13), gpu_reference (the reference's
                     eager iteration on an upstream-structure CUDA rasterizer, same box), cpu_baseline, gnn / gnn_train.
reference          : the reference's iteration on the host cores (oracle port: C rasterizer restatement + CPU PyTorch).
upstream_structure : the reference's eager iteration on this GPU (baseline/upstream_structure) — the "reference CUDA
                     rasterizer on 1 GPU" baseline of the north-star, as a clearly-labelled structural re-creation.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

HBM_FALLBACK_GBS = 6650.0


def make_config(G, world, n_cams=4):
    """The SAME dict for every arm (ours / reference / upstream_structure): the driver compares them."""
    return {"workload": "train_gs.py steady-state iteration (t>0), %dk Gaussians, 4 cams @640x480, 1 episode per GPU" % (G // 1000),
            "gaussians": G, "cameras": n_cams, "image": "640x480", "knn": 20,
            "parallelism": "episode-per-gpu x%d" % world,
            "l2": "GPU arms: 256 MiB L2 flush between timed steps (excluded from step time)",
            "timing": "GPU arms: CUDA events per step on the launch stream, max over ranks; CPU arm: wall clock"}


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  nvidia-smi needs ~a second before its first line, so
    the process is started before the warm-up; mark() / stop() bracket the timed region and only the lines read between them
    are used (if the region was shorter than one sampling period, the closest line on either side is used and flagged)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines, self.t0 = gpu_index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        t = time.time()
        while self.proc and not self.lines and time.time() - t < timeout:
            time.sleep(0.05)

    def mark(self):
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        t1 = time.time()
        time.sleep(0.06)   # let the line that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.06]
        note = None
        if not inside and self.lines:
            inside = [min(self.lines, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[1]]
            note = "timed region shorter than one sampling period: closest sample"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------------------------
def build_gpu_problem(G, seed, device):
    from gs_dynamics_b200 import workloads
    return workloads.tracking_problem_gpu(G, seed, device)


def time_blend_backward(params, dataset, capacity, flush, iters=10):
    """CUDA-event duration of (a) the dominant kernel (blend backward, 6 channels, geometry-only) alone and (b) the whole
    rasterizer forward + backward of one camera, L2 flushed before each measurement."""
    import ctypes as C
    from gs_dynamics_b200 import tracking as TR, rasterizer as R, _lib
    data = dataset[0]
    with torch.no_grad():
        rv = TR.params2rendervar(params)
        seg = params["seg_colors"].detach()
        args = (data["cam"], rv["means3D"], rv["opacities"], rv["colors_precomp"], rv["scales"], rv["rotations"])
        color, radii, depth, st = R.raster_forward(*args, colors1=seg, capacity=capacity)
        dL = torch.randn_like(color)
        sz = R._workspace_bytes(st.G, st.W, st.H, st.n_sets, st.capacity)
        partial = torch.empty(sz[3], dtype=torch.uint8, device=color.device)
        b = _lib.GsdRasterBwd()
        b.fwd = st.desc
        b.dL_dcolor, b.partial_ws = dL.data_ptr(), partial.data_ptr()  # no colour/opacity outputs: the steady-state (geometry-only) kernel
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts, tr = [], []
        for i in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.lib().gsd_raster_backward_stage(C.byref(b), 1, stream), "gsd_raster_backward_stage")
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        for i in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c2, _, _, st2 = R.raster_forward(*args, colors1=seg, capacity=capacity)
            R.raster_backward(st2, dL, need_means2D=False, geom_only=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                tr.append(e0.elapsed_time(e1) * 1e-3)
        R_inst = int(st.status[0].item())
    return float(np.mean(ts)), float(np.mean(tr)), R_inst


def load_profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def time_device_steps(step_obj, n_cams, rng, steps, warmup, flush, barrier=None):
    for _ in range(warmup):
        step_obj.step(rng.randrange(n_cams))
    if barrier:
        barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_obj.step(rng.randrange(n_cams))
        e1.record()
        evs.append((e0, e1))
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) * 1e-3


def run_episode(device, G, frames, iters, seed=0):
    """BASELINE configs[2] (train_gs.py:19-46 for t > 0): per frame — target images uploaded from pinned host memory,
    initialize_per_timestep (constant-velocity warm start, Adam reset, prev_* tables, edge-record pack), exactly `iters`
    iterations through the per-camera graphs captured ONCE, one capacity check.  Targets drift by 0.5 mm per frame
    (SURVEY.md §8d config 3).  Everything between the first frame's start and the last frame's end is inside the timed region
    (wall clock with a device sync at both ends); graph capture happens before it and is reported separately."""
    from gs_dynamics_b200 import tracking as TR, workloads, rasterizer as R
    prob = workloads.tracking_problem(G, seed)
    params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, seed, device, prob=prob)
    tgt = {k: t.to(device) for k, t in prob["target"].items()}
    ones = torch.ones_like(tgt["colors_precomp"])
    host_frames = []
    with torch.no_grad():
        for f in range(frames):   # the "dataset on disk": per-frame images in pinned host memory
            shift = torch.tensor([0.0005 * (f + 1), 0.0, 0.0], device=device)
            fr = []
            for d in dataset:
                out, _, _, _ = R.raster_forward(d["cam"], tgt["means3D"] + shift, tgt["opacities"], tgt["colors_precomp"], tgt["scales"],
                                                tgt["rotations"], colors1=ones)
                fr.append((out[:3].cpu().pin_memory(), workloads.seg_target_from_mask(out[3]).cpu().pin_memory()))
            host_frames.append(fr)
    rng = random.Random(seed)
    t0 = time.time()
    params, variables = TR.initialize_per_timestep(params, variables, opt)
    step = TR.FusedTrackingStep(params, variables, opt, dataset)
    step.prepare()
    torch.cuda.synchronize()
    t_capture = time.time() - t0
    hw_max, recaptures = 0, 0
    t1 = time.time()
    for f in range(frames):
        if f > 0:
            params, variables = TR.initialize_per_timestep(params, variables, opt)
        for c, (im, seg) in enumerate(host_frames[f]):
            step.set_target(c, im, seg)
        cap0 = dict(step.capacity)
        hw = TR.run_frame(step, [rng.randint(0, len(dataset) - 1) for _ in range(iters)])
        hw_max = max(hw_max, hw)
        recaptures += int(cap0 != step.capacity)
    torch.cuda.synchronize()
    dt = time.time() - t1
    h2d = sum(im.numel() * 4 + seg.numel() * 4 for im, seg in host_frames[0])
    return {"metric": "tracked frames/sec (2 000 iterations each)", "value": frames / dt, "unit": "frames/s", "iters_per_s": frames * iters / dt,
            "frames": frames, "iters_per_frame": iters, "seconds": dt, "graph_capture_seconds_once": t_capture, "recaptures": recaptures,
            "max_instances_seen": hw_max, "capacity": min(step.capacity.values()), "h2d_bytes_per_frame": h2d,
            "config": {"workload": "train_gs.py episode, frames t>0: %dk Gaussians, 4 cams @640x480, per-frame target upload + "
                                   "initialize_per_timestep + %d iterations + 1 capacity check; graphs captured once" % (G // 1000, iters)}}


def run_render_720p(device, G, iters=20):
    """A13: Renderer.render + the ones-colour mask render of predict.py:116-123 at 1280x720 (3 600 tiles), per frame."""
    from gs_dynamics_b200 import render as RD, scenes
    W0, H0, cams = scenes.demo_cameras()
    k, w2c = cams[0]
    k = k.copy(); k[0] *= 1280 / W0; k[1] *= 720 / H0
    act = {kk: v.to(device) for kk, v in scenes.activate(scenes.synthetic_scene(G, 0)).items()}
    act["means2D"] = torch.zeros_like(act["means3D"])
    r = RD.Renderer(device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.render(w2c, k, act)
        r.render_mask(w2c, k, act)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return {"metric": "predict.py render leg: image + mask render per frame", "ms_per_frame": float(np.median(ts)),
            "value": 1e3 / float(np.median(ts)), "unit": "frames/s", "config": {"workload": "Renderer.render + render_mask, %dk Gaussians, 1280x720 (incl. the reference's exact-size num_rendered sync per call)" % (G // 1000)}}


def run_upstream_structure(G, steps, warmup, device, seed=0):
    """The reference's eager iteration on this GPU (baseline/upstream_structure).  Device time per step incl. its host syncs
    (CUDA events around each iteration, L2 flushed between iterations like the other GPU arm)."""
    from baseline import upstream_structure as U
    from gs_dynamics_b200 import workloads
    prob = workloads.tracking_problem(G, seed)
    _, _, _, dataset, _ = workloads.tracking_problem_gpu(G, seed, device, prob=prob)   # same targets / cameras as our arm
    params = {k: torch.nn.Parameter(v.to(device).contiguous()) for k, v in prob["params"].items()}
    params["rgb_colors"].requires_grad = False
    variables = {k: (t.to(device) if isinstance(t, torch.Tensor) else t) for k, t in prob["variables"].items()}
    opt = U.make_optimizer(params, variables["scene_radius"])
    Ras, label = U.rasterizer_module(prefer_real=True)
    rng = random.Random(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    for _ in range(warmup):
        U.iteration(Ras, params, dataset[rng.randrange(len(dataset))], variables, opt)
    torch.cuda.synchronize()
    total, t0 = 0.0, time.time()
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = U.iteration(Ras, params, dataset[rng.randrange(len(dataset))], variables, opt)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1) * 1e-3
    return {"value": steps / total, "unit": "iters/s", "ms_per_step": 1e3 * total / steps, "steps": steps, "warmup": warmup,
            "rasterizer": label, "final_loss": float(loss), "wall_s": time.time() - t0,
            "kind": "reference's eager get_loss + backward + torch.optim.Adam (train_utils.py:167-246, train_gs.py:31-39) on 1 GPU"}


def run_ours(args):
    import torch.distributed as dist
    from gs_dynamics_b200 import tracking as TR, _lib, dist as gdist
    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    gdist.init(backend="nccl", device=device)  # NCCL: barrier + max-over-ranks timing only; episodes share nothing
    G = args.gaussians
    params, variables, opt, dataset, host = build_gpu_problem(G, seed=rank, device=device)
    n_cams = len(dataset)
    rng = random.Random(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    step_obj = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=True)
    step_obj.prepare()
    # launches of ONE iteration: counted on an eager run of the same code (graph replays are not seen by the host counter)
    eager = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=False)
    eager.capacity = dict(step_obj.capacity)
    snap = eager.snapshot()
    a0 = _lib.launch_count()
    eager.step(0)
    a1 = _lib.launch_count()
    eager.restore(snap)
    own_per_iter, lib_per_iter = a1[0] - a0[0], a1[1] - a0[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_obj.step(rng.randrange(n_cams))
    if rank == 0:
        sampler.wait_first()
    barrier()
    sampler.mark()
    t_wall0 = time.time()
    dev_s = time_device_steps(step_obj, n_cams, rng, args.steps, 0, flush, barrier)
    t_wall = time.time() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    hw, n_over = step_obj.check_capacity(raise_on_overflow=True)   # one 8-byte read: no replay dropped instances

    # ---- end to end through the public API with host buffers: every step its camera image + seg are copied from pinned host
    # memory into the step's target buffers (plus their SSIM window statistics), and the loss is read back and consumed on the
    # host.  The copy for step i+1 is issued on a side stream while step i computes (double buffering per camera; when the next
    # step uses the same camera the copy waits for the running step).  All of it is inside the timed region.
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    # the step's inputs as the reference's dataset holds them at rest (train_utils.py:66-75: 8-bit PNGs): RGB image [H,W,3] uint8 +
    # segmentation mask [H,W] uint8 in pinned host memory; the float planes (im / 255 | seg, 0, 1 - seg) are built on the device
    host_u8 = [((im.permute(1, 2, 0) * 255.0).round().clamp_(0, 255).to(torch.uint8).contiguous().pin_memory(),
                seg[0].round().clamp_(0, 1).to(torch.uint8).contiguous().pin_memory()) for im, seg in host]
    h2d = host_u8[0][0].numel() + host_u8[0][1].numel()
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream()

    def prefetch(c, after=None):
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)
            step_obj.set_target_u8(c, host_u8[c][0], host_u8[c][1])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def run_e2e(n_steps, timed):
        seq = [rng.randrange(n_cams) for _ in range(n_steps + 1)]
        total = 0.0
        last = None
        copy_ev = prefetch(seq[0])
        torch.cuda.synchronize()
        for i in range(n_steps):
            if timed:
                flush.zero_()                              # stream-ordered before e0: no host sync needed (and none wanted —
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # it would starve the GPU)
            e0.record()
            l = step_obj.step(seq[i])                      # inputs of this step landed inside the previous timed region
            done = torch.cuda.Event()
            done.record(main)
            # next step's inputs: overlap with this step unless it reads the same camera buffers; the timed region of this
            # step ends only when that copy has landed too, so every H2D byte is paid for inside a timed region
            copy_ev = prefetch(seq[i + 1], after=done if seq[i + 1] == seq[i] else None)
            main.wait_event(copy_ev)
            loss_host.copy_(l, non_blocking=True)
            e1.record()
            main.synchronize()                             # the caller consumes the loss every step
            last = float(loss_host)
            if timed:
                total += e0.elapsed_time(e1) * 1e-3
        torch.cuda.synchronize()
        return total, last

    run_e2e(max(3, args.warmup // 4), False)
    barrier()
    e2e_s, last_loss = run_e2e(args.steps, True)
    barrier()

    # ---- rooflines
    peak, peak_src = measured_peaks()
    t_blend, t_raster, R_inst = time_blend_backward(params, dataset, step_obj.capacity[0], flush)
    P = 640 * 480
    alg_bytes = 56.0 * R_inst + 32.0 * P + 20.0 * G  # blend backward, 6 channels, geometry-only partials, per launch
    achieved = alg_bytes / t_blend / 1e9
    prof = load_profile_json("roofline_traffic.json") or {}
    pk = prof.get(str(G), {}) if isinstance(prof.get(str(G)), dict) else {}
    traffic = pk.get("blend_backward_dram_bytes_per_launch", prof.get("blend_backward_dram_bytes_per_launch") if G == 50000 else None)
    warp_insts = pk.get("blend_backward_warp_instructions_per_launch")
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6            # warp instructions / s: 148 SMs x 4 schedulers x 1 per clock
    agg_bytes = 260.0 * G + 124.0 * R_inst + 44.0 * P   # one 3-channel fwd+bwd of one camera, every byte once
    roofline = {"bound": "hbm", "kernel": "gsd_blend_bwd_chunk_kernel<6,geom> (+ prefix kernel)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                "kernel_us": t_blend * 1e6, "instances_per_camera": R_inst,
                "issue_slot": {"note": "the blend kernels are instruction-issue bound, not HBM bound (DESIGN.md §6): this is the binding roof",
                               "warp_instructions_per_launch": warp_insts, "source": "profiles/roofline_traffic.json (ncu smsp__inst_executed.sum)" if warp_insts else None,
                               "peak_warp_inst_per_s": issue_peak, "sm_mhz": sm_mhz,
                               "frac": (warp_insts / t_blend / issue_peak) if warp_insts else None},
                "aggregate": {"what": "SURVEY.md §8(d) rasterizer unit (260 G + 124 R + 44 P bytes: one 3-channel forward+backward of one camera) "
                                      "over the time of OUR fused 6-channel forward+backward (which replaces TWO such reference passes)",
                              "bytes": agg_bytes, "raster_fwd_bwd_us": t_raster * 1e6, "achieved": agg_bytes / t_raster / 1e9,
                              "frac": agg_bytes / t_raster / 1e9 / peak, "frac_counting_both_replaced_passes": 2 * agg_bytes / t_raster / 1e9 / peak,
                              "whole_step_frac": 2 * agg_bytes / (dev_s / args.steps) / 1e9 / peak},
                "stages": load_profile_json("r2_stage_table_%dk.json" % (G // 1000))}

    times = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(times[0]), float(times[1])
    line = None
    if rank == 0:
        line = {
            "metric": "tracking iters/sec", "value": world * args.steps / dev_s, "unit": "iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(G, world, n_cams),
            "e2e": {"value": world * args.steps / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(own_per_iter * args.steps), "gpu_launches_per_step": int(own_per_iter),
            "library_launches_per_step": int(lib_per_iter),
            "roofline": roofline,
            "capacity_check": {"max_instances_seen": hw, "overflowed_calls": n_over, "capacity": min(step_obj.capacity.values())},
            "clocks": clocks, "wall_s": t_wall, "final_loss": last_loss,
        }
    del step_obj, eager
    return line, rank, world


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port on the host cores)
# ------------------------------------------------------------------------------------------------------------------
def build_cpu_problem(G, seed):
    from gs_dynamics_b200 import workloads
    from oracle import raster_c, tracking_cpu
    prob = workloads.tracking_problem(G, seed)
    params = {k: torch.nn.Parameter(v.clone().contiguous()) for k, v in prob["params"].items()}
    params["rgb_colors"].requires_grad = False
    variables = dict(prob["variables"])
    tgt = prob["target"]
    dataset = []
    for c in prob["cams"]:
        m = c["mats"]
        args = (tgt["means3D"], tgt["colors_precomp"], tgt["opacities"], tgt["scales"], tgt["rotations"], m["viewmatrix"],
                m["projmatrix"], torch.zeros(3), m["tanfovx"], m["tanfovy"], m["image_height"], m["image_width"])
        im = torch.from_numpy(raster_c.forward(*args)["color"])
        a2 = list(args); a2[1] = torch.ones_like(tgt["colors_precomp"])
        mask = torch.from_numpy(raster_c.forward(*a2)["color"][0])
        from gs_dynamics_b200.workloads import seg_target_from_mask
        dataset.append({"cam": m, "im": im, "seg": seg_target_from_mask(mask), "id": c["id"]})
    opt = tracking_cpu.make_optimizer(params, variables["scene_radius"])
    return params, variables, opt, dataset


def cpu_iterations(G, steps, warmup, budget_s):
    from oracle import tracking_cpu, raster_c
    torch.set_num_threads(os.cpu_count() or 1)
    raster_c.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1 to its workers
    params, variables, opt, dataset = build_cpu_problem(G, 0)
    rng = random.Random(0)
    for _ in range(warmup):
        tracking_cpu.iteration(params, dataset[rng.randrange(len(dataset))], variables, opt)
    done, t0 = 0, time.time()
    while done < steps and (done == 0 or time.time() - t0 < budget_s):
        tracking_cpu.iteration(params, dataset[rng.randrange(len(dataset))], variables, opt)
        done += 1
    dt = time.time() - t0
    return done, dt, max(raster_c.num_threads(), torch.get_num_threads())


def run_reference(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return None, rank, world
    G = args.gaussians
    done, dt, cores = cpu_iterations(G, args.steps, args.warmup, budget_s=150.0)
    val = done / dt
    sample = "%d of the requested %d iterations (time-bounded), same config as the GPU arm" % (done, args.steps)
    line = {"impl": "reference", "metric": "tracking iters/sec", "value": val, "unit": "iters/s", "n_gpus": world, "steps": done,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(G, world),
            "note": "reference iteration on host cores: oracle port (C rasterizer restatement + CPU PyTorch); the reference's own "
                    "rasterizer is CUDA-only and un-vendored.  A CPU process does not scale with --gpus: rank 0 alone runs it, so at "
                    "N > 1 the driver's ratio compares N GPUs with this one process",
            "cpu_baseline": {"value": val, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line, rank, world


def run_upstream_arm(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return None, rank, world
    torch.cuda.set_device(local_rank)
    r = run_upstream_structure(args.gaussians, min(args.steps, 200), min(args.warmup, 10), torch.device("cuda", local_rank))
    line = {"impl": "upstream_structure", "metric": "tracking iters/sec", "value": r["value"], "unit": "iters/s", "n_gpus": 1, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": make_config(args.gaussians, world), "gpu_reference": r,
            "e2e": {"value": r["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line, rank, world


# ------------------------------------------------------------------------------------------------------------------
# secondary metric: GNN rollout steps/sec (BASELINE.json configs[3]: sloth cfg, 2k particles, 8-NN, 50-step horizon)
# ------------------------------------------------------------------------------------------------------------------
def run_gnn_ours(device, steps=50, warmup=5, n_obj=2000, seed=1):
    from gs_dynamics_b200 import gnn, workloads as GO
    cfg = GO.sloth_cfg(512)
    model = gnn.DynamicsPredictor(dict(cfg), device).to(device).eval()
    model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=1e-3))
    gi = GO.make_graph_inputs(n_obj, seed, "sloth")
    p0, eef = gi["state"][0, :, :n_obj].to(device), gi["state"][0, :, n_obj:].to(device)
    ro = gnn.GnnRollout(model, p0, eef, 0.075, 8, True, use_graph=True)
    delta = torch.tensor([0.005, 0.0, 0.0], device=device)
    for _ in range(warmup):
        ro.step(delta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ro.step(delta)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    # end to end: tool displacement comes from the host every step, predicted particle positions are read back
    host_delta = torch.tensor([0.005, 0.0, 0.0]).pin_memory()
    host_pred = torch.empty((1, n_obj, 3)).pin_memory()
    t0 = time.time()
    for _ in range(steps):
        ro.eef_delta.copy_(host_delta, non_blocking=True)
        pred = ro.step()
        host_pred.copy_(pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    dt_e2e = time.time() - t0
    skin = time_skinning(device, model, ro)
    roof = gnn_roofline(device, model, ro, dt / steps)
    return {"roofline": roof, "skinning": skin, "metric": "GNN steps/sec", "value": steps / dt, "unit": "steps/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "config": {"workload": "predict.py GNN rollout step (edge build + forward + history shift), sloth cfg nf=512, "
                                   "%d particles + 1 tool, 8-NN, connect_all, %d-step horizon, random-init weights with the motion head scaled 1e-3 so the cloud keeps its 8-NN graph" % (n_obj, steps)},
            "e2e": {"value": steps / dt_e2e, "unit": "steps/s", "h2d_bytes_per_step": 12, "d2h_bytes_per_step": n_obj * 12}}


def gnn_roofline(device, model, ro, step_s, iters=20):
    """(i) the north-star's scatter-reduce (gsd_gnn_aggregate) timed alone against the HBM roof: algorithmic bytes
    4 E F (A_e) + 8 E (indices) + 12 N F (P_r | P_s read, agg written); (ii) the step's real FLOPs (weight-split
    formulation, 3 error-compensated TF32 products each) against the dense TF32 tensor peak (= half the measured bf16 peak)."""
    from gs_dynamics_b200 import gnn
    peak, src = measured_peaks()
    edges = gnn.construct_edges_index(ro.states[:, -1], ro.adj_thresh, ro.state_mask, ro.eef_mask, topk=ro.topk, connect_all=ro.connect_all, n_tool=1)
    E, N, Fd = int(edges.n_edges[0]), edges.N, 512
    A = torch.randn(1, edges.capacity, Fd, device=device)
    P = torch.randn(N, 2 * Fd, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    ts = []
    with torch.no_grad():
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gnn._Aggregate.apply(A, P, edges)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(ts))
    alg = 4.0 * E * Fd + 8.0 * E + 12.0 * N * Fd
    in_dim = 14
    flops = N * 2 * (in_dim * Fd + 2 * Fd * Fd) + E * 2 * (14 * Fd + 2 * Fd * Fd) + E * 2 * Fd * Fd + 3 * (N * 2 * Fd * 2 * Fd + N * 2 * 2 * Fd * Fd) + N * 2 * (2 * Fd * Fd + 3 * Fd)
    tf32_peak = None
    pj = load_profile_json("../MEASURED_PEAKS.json")
    if pj and "bf16_tflops" in pj:
        tf32_peak = 0.5 * float(pj["bf16_tflops"])
    return {"bound": "hbm", "kernel": "gsd_gnn_aggregate (rows + heavy tool row)", "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s",
            "frac": alg / t / 1e9 / peak, "kernel_us": t * 1e6, "algorithmic_bytes": alg, "edges": E, "nodes": N, "peak_source": src,
            "note": "flushed L2: in the step the operands are L2-resident (written by the preceding GEMM), so the in-step time is lower",
            "dense_layers": {"real_flops_per_step": flops, "x3_compensated_tf32_flops": 3 * flops, "achieved_tflops_whole_step": 3 * flops / step_s / 1e12,
                             "tf32_peak_tflops": tf32_peak, "frac_whole_step": (3 * flops / step_s / 1e12 / tf32_peak) if tf32_peak else None,
                             "peak_source": "0.5 x MEASURED_PEAKS.json bf16_tflops (dense TF32 = half of bf16)"}}


def run_gnn_batched(device, B=1000, n_obj=100, steps=10, warmup=3):
    """MPPI-style batched rollout (plan.py:25-154: bsz = 1000 perturbed action samples, ~100 particles + tool, graph rebuilt
    every step): sample-steps per second of one CUDA-graph step over the whole batch."""
    from gs_dynamics_b200 import gnn, workloads as GO
    cfg = GO.sloth_cfg(512)
    model = gnn.DynamicsPredictor(dict(cfg), device).to(device).eval()
    model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=1e-3))
    gi = GO.make_graph_inputs(n_obj, 7, "sloth")
    p0, eef = gi["state"][0, :, :n_obj].to(device), gi["state"][0, :, n_obj:].to(device)
    ro = gnn.GnnRollout(model, p0, eef, 0.075, 5, True, use_graph=True, batch=B)
    g = torch.Generator(device="cpu").manual_seed(0)
    deltas = (0.005 * torch.randn(B, 3, generator=g)).to(device)
    deltas[:, 2] = 0
    for _ in range(warmup):
        ro.step(deltas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ro.step(deltas)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    return {"metric": "MPPI batched rollout sample-steps/sec", "value": B * steps / dt, "unit": "sample-steps/s", "ms_per_step": 1e3 * dt / steps,
            "config": {"workload": "plan.py dynamics(): %d action samples x (%d particles + tool), topk 5, connect_all, nf 512, graph rebuilt every step" % (B, n_obj)}}


def time_skinning(device, model, ro, n_gauss=100000, iters=20):
    """The step after the GNN (interpolate_motions): 100k Gaussians skinned from the 2000 particles of the
    rollout graph.  Device time per call, L2 flushed between calls; CPU: the oracle restatement on a 5k-Gaussian sample."""
    from gs_dynamics_b200 import gnn, skinning as SK
    g = torch.Generator().manual_seed(0)
    xyz = (torch.rand(n_gauss, 3, generator=g) * torch.tensor([0.5, 0.5, 0.1])).to(device)
    quat = torch.nn.functional.normalize(torch.randn(n_gauss, 4, generator=g), dim=-1).to(device)
    bones = ro.states[0, -1, :ro.nobj].clone()
    motions = torch.randn(ro.nobj, 3, generator=g).to(device) * 0.003
    edges = gnn.construct_edges_index(ro.states[:, -1], ro.adj_thresh, ro.state_mask, ro.eef_mask, topk=ro.topk, connect_all=ro.connect_all, n_tool=1)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        SK.interpolate_motions(bones, motions, edges, xyz, quat=quat, return_weights=False)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    out = {"us_per_call": float(np.median(ts)), "gaussians": n_gauss, "bones": int(ro.nobj),
           "pairs_per_s": n_gauss * ro.nobj / (float(np.median(ts)) * 1e-6)}
    try:
        from oracle import skinning_oracle as SO
        Rr, Rs = edges.dense()
        rel = SK.relations_to_matrix(Rr, Rs)[:ro.nobj, :ro.nobj].cpu()
        sample = 5000
        t0 = time.time()
        SO.interpolate_motions(bones.cpu(), motions.cpu(), rel, xyz[:sample].cpu(), quat=quat[:sample].cpu(), dtype=torch.float32)
        out["cpu_baseline"] = {"seconds": time.time() - t0, "kind": "port", "cores": torch.get_num_threads(),
                               "sample": "%d Gaussians x %d bones (reference formulation: per-bone Python loop + dense weights)" % (sample, ro.nobj)}
    except Exception as ex:  # the CPU leg is a reported baseline only
        out["cpu_baseline"] = {"error": repr(ex)}
    return out


def run_gnn_cpu(budget_s=8.0, n_obj=2000, seed=1):
    """The reference formulation (dense one-hot bmm, oracle/gnn_oracle.py) on the host cores, bounded sample."""
    from oracle import gnn_oracle as GO
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = GO.sloth_cfg(512)
    sd = GO.make_state_dict(cfg, 0, head_scale=1e-3)
    gi = GO.make_graph_inputs(n_obj, seed, "sloth")
    done, t0 = 0, time.time()
    with torch.no_grad():
        while done == 0 or time.time() - t0 < budget_s:
            recv, send = GO.construct_edges(gi["state"][0, -1], 0.075, gi["state_mask"], gi["eef_mask"], 8, True)
            Rr, Rs = GO.one_hot_edges(recv, send, n_obj + 1)
            GO.forward(sd, cfg, gi["state"], gi["attrs"], Rr[None], Rs[None], gi["p_instance"], gi["action"])
            done += 1
    dt = time.time() - t0
    return {"value": done / dt, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps in %.1f s (dense one-hot reference formulation on host cores)" % (done, dt)}


def run_gnn_train(device, world, steps=20, warmup=3, B=64, n_obj=100, n_future=5):
    """GNN training iteration (train.py:176-218 at the reference's batch_size 64 / n_future 5, nf = 512,
    ~100 object particles, topk 5): zero_grad + 5-step unroll + backward + gradient-bucket all-reduce (N > 1, NCCL AVG) + Adam,
    captured in ONE CUDA graph per rank (gnn_train.GraphedTrainStep).  Every rank trains on its own batch of a shared, learnable
    mapping (weak scaling); device time per step, max over ranks; the reported losses are means over ALL ranks."""
    import torch.distributed as dist
    from gs_dynamics_b200 import gnn, gnn_train, workloads as GO
    rank = int(os.environ.get("RANK", "0"))
    cfg = GO.sloth_cfg(512)
    model = gnn.DynamicsPredictor(dict(cfg), device).to(device).train()
    model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=0.05))
    batch = {k: v.to(device) for k, v in GO.make_training_batch(B, n_obj, 100 + rank, "sloth", n_future, learnable=True).items()}
    batch["Rr"] = gnn.construct_edges_index(batch["state"][:, -1], 0.075, batch["state_mask"], batch["eef_mask"], topk=5, connect_all=True)
    batch["Rs"] = None
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)    # train.py's learning rate
    bucket = gnn_train.GradientBucket(model.parameters())
    funcs = gnn_train.default_loss_funcs({"mse_loss": 1.0, "length_loss": 0.05})

    def mean_loss(l):
        t = l.detach().double().reshape(1).clone()
        if world > 1:
            dist.all_reduce(t)
        return float(t[0]) / world
    # eager iteration time (the round-1 path) for reference, then the graphed step
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l_first = gnn_train.train_iteration(model, opt, batch, n_future, funcs, bucket)[0]
    loss_first = mean_loss(l_first)
    e0.record()
    for _ in range(3):
        gnn_train.train_iteration(model, opt, batch, n_future, funcs, bucket)
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / 3
    step = gnn_train.GraphedTrainStep(model, opt, batch, n_future, funcs, bucket, warmup=warmup)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    E = int(batch["Rr"].n_edges.sum())
    return {"metric": "GNN training samples/sec", "value": world * B * steps / dt, "unit": "samples/s", "ms_per_iteration": 1e3 * dt / steps,
            "eager_ms_per_iteration": eager_ms, "steps": steps, "n_gpus": world,
            "allreduce_bytes_per_step": int(bucket.flat.numel() * 4) if world > 1 else 0, "allreduce_op": "NCCL AVG, captured in the step's CUDA graph",
            "loss_first_mean_over_ranks": loss_first, "loss_last_mean_over_ranks": mean_loss(loss), "iterations_between": 4 + warmup + 1 + steps,
            "config": {"workload": "train.py iteration: batch %d x (%d particles + pad + tool), %d edges/batch, n_future %d, nf 512, "
                                   "mse + 0.05 length loss, Adam lr 1e-4; DP = one batch per GPU + one flat-bucket all-reduce; whole iteration in one CUDA graph" % (B, n_obj, E, n_future)}}


def run_gnn_train_cpu(B=4, n_obj=100, n_future=5):
    """The reference formulation (dense one-hot bmm, autograd) of the same iteration on the host cores: one fwd+bwd of a
    B = 4 sample of the batch (the oracle is a pure function of the state_dict; Adam is negligible)."""
    from oracle import gnn_oracle as GO
    torch.set_num_threads(os.cpu_count())
    cfg = GO.sloth_cfg(512)
    sd = {k: v.clone().requires_grad_(True) for k, v in GO.make_state_dict(cfg, 0, head_scale=0.05).items()}
    batch = GO.make_training_batch(B, n_obj, 100, "sloth", n_future)
    Rr, Rs = GO.batch_edges(batch, 0.075, 5, True)
    t0 = time.time()
    loss, _ = GO.unrolled_loss(sd, cfg, batch, Rr, Rs, n_future, 1.0, 0.05)
    loss.backward()
    dt = time.time() - t0
    return {"value": B / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "1 forward+backward of a %d-sample batch in %.1f s (dense one-hot reference formulation, autograd, host cores)" % (B, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)  # one reference frame = 2000 iterations (train_gs.py:25)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "upstream_structure"])
    ap.add_argument("--gaussians", type=int, default=100000)   # BASELINE configs[2] / the north-star's target size
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--episode-frames", type=int, default=20)
    ap.add_argument("--no-gnn", action="store_true", help="skip the secondary GNN measurements")
    ap.add_argument("--no-extras", action="store_true", help="skip episode / 50k / render / gpu_reference side objects")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl in ("reference", "upstream_structure"):
        line, rank, world = run_reference(args) if args.impl == "reference" else run_upstream_arm(args)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback in the product path (use --impl reference for the CPU arm)")
    line, rank, world = run_ours(args)
    gtrain = None
    if not args.no_gnn:
        try:   # every rank takes part (N > 1: the gradient all-reduce is a real NCCL collective)
            gtrain = run_gnn_train(torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))), world)
        except Exception as ex:
            gtrain = {"error": repr(ex)}
    if rank == 0:
        if gtrain is not None:
            line["gnn_train"] = gtrain
        if world == 1:
            dev0 = torch.device("cuda", 0)

            def side(name, fn):   # a side object must never take the headline line down
                try:
                    line[name] = fn()
                except Exception as ex:
                    line[name] = {"error": repr(ex)}
            if not args.no_extras:
                side("gpu_reference", lambda: run_upstream_structure(args.gaussians, 100, 5, dev0))
                if "value" in line.get("gpu_reference", {}):
                    line["gpu_reference"]["ours_over_reference_e2e"] = line["e2e"]["value"] / line["gpu_reference"]["value"]
                    line["gpu_reference"]["ours_over_reference_device"] = line["value"] / line["gpu_reference"]["value"]
                side("episode", lambda: run_episode(dev0, args.gaussians, args.episode_frames, 2000))
                side("secondary_50k", lambda: secondary_50k(dev0))
                side("render_1280x720", lambda: run_render_720p(dev0, args.gaussians))
            done, dt, cores = cpu_iterations(args.gaussians, 10 ** 9, 1, budget_s=args.cpu_baseline_seconds)
            line["cpu_baseline"] = {"value": done / dt, "unit": "iters/s", "cores": cores, "kind": "port",
                                    "sample": "%d iterations in %.1f s of the same workload (oracle port on host cores)" % (done, dt)}
            if not args.no_gnn:
                try:
                    g = run_gnn_ours(dev0)
                    g["cpu_baseline"] = run_gnn_cpu()
                    g["mppi_batch"] = run_gnn_batched(dev0)
                    line["gnn"] = g
                    if "error" not in line.get("gnn_train", {"error": 1}):
                        line["gnn_train"]["cpu_baseline"] = run_gnn_train_cpu()
                except Exception as ex:  # the secondary metric must never take the headline line down
                    line["gnn"] = {"error": repr(ex)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def secondary_50k(device, steps=500, warmup=20):
    """BASELINE configs[1] (single frame, 50k Gaussians): device-resident iterations/s, same method as the headline."""
    from gs_dynamics_b200 import tracking as TR
    params, variables, opt, dataset, host = build_gpu_problem(50000, 0, device)
    step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=True)
    step.prepare()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    dt = time_device_steps(step, len(dataset), random.Random(0), steps, warmup, flush)
    step.check_capacity()
    return {"metric": "tracking iters/sec", "value": steps / dt, "unit": "iters/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "config": make_config(50000, 1)}


if __name__ == "__main__":
    main()
