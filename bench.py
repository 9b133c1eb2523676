#!/usr/bin/env python
"""bench.py - throughput of the raster / raster_pullback! hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA library
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port), host cores

A "step" is one pass of the hot path over one pose batch: forward splat + pullback of the workload
BASELINE.json's metric is quoted on (configs[1]: 3d->2d, 100k points x 4096 poses, 256x256, Float32), with the
inputs resident in HBM.  With N > 1 ranks (torchrun) every rank owns a 4096-pose shard (weak scaling), the points
are replicated, and the step ends with the single all-reduce of [d_points; d_point_weight].
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: n_in, n_out, P, B (per GPU), grid, dtype, explicit weights/background, what to run
    "cfg1": dict(n_in=3, n_out=2, P=10_000, B=64, grid=(128, 128), dtype="f64", weights=False, ops="fwd+bwd",
                 label="3d->2d, 10k points x 64 poses, 128x128, Float64 (README.md:189)"),
    "cfg2": dict(n_in=3, n_out=2, P=100_000, B=4096, grid=(256, 256), dtype="f32", weights=False, ops="fwd+bwd",
                 label="3d->2d, 100k points x 4096 poses, 256x256, Float32 (cryo-EM projection batch)"),
    "cfg3": dict(n_in=3, n_out=3, P=1_000_000, B=16, grid=(256, 256, 256), dtype="f32", weights=False, ops="fwd+bwd",
                 label="3d->3d voxelisation, 1M points x 16 poses, 256^3, Float32"),
    "cfg4": dict(n_in=2, n_out=2, P=1_000_000, B=1024, grid=(512, 512), dtype="f32", weights=True, ops="fwd+bwd",
                 label="2d->2d, 1M points x 1024 poses, 512x512, Float32, point weights + background"),
    "cfg5": dict(n_in=3, n_out=2, P=1_000_000, B=2048, grid=(128, 128), dtype="f32", weights=False, ops="bwd",
                 label="3d->2d pullback-only, 1M points x 2048 poses per GPU (16384 over 8 GPUs), 128x128, Float32"),
    # the other rows of the reference's README "Timings" table (README.md:189-193; element type not stated: Float64 here)
    "readme2": dict(n_in=3, n_out=2, P=10_000, B=64, grid=(1024, 1024), dtype="f64", weights=False, ops="fwd+bwd",
                    label="README.md:190: 3d->2d, 10k points x 64 images, 1024x1024, Float64"),
    "readme3": dict(n_in=3, n_out=2, P=100_000, B=64, grid=(128, 128), dtype="f64", weights=False, ops="fwd+bwd",
                    label="README.md:191: 3d->2d, 100k points x 64 images, 128x128, Float64"),
    "readme4": dict(n_in=3, n_out=2, P=100_000, B=64, grid=(1024, 1024), dtype="f64", weights=False, ops="fwd+bwd",
                    label="README.md:192: 3d->2d, 100k points x 64 images, 1024x1024, Float64"),
    "readme5": dict(n_in=3, n_out=3, P=100_000, B=1, grid=(1024, 1024, 1024), dtype="f64", weights=False, ops="fwd+bwd",
                    label="README.md:193: 3d->3d, 100k points x 1 image, 1024^3, Float64"),
}
_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved
    descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


METRIC = "splats/sec (points x poses), forward + pullback"
UNIT = "splats/s"


def synth_inputs(cfg, seed, rank):
    """Synthetic inputs with the reference fixtures' distributions (test/data.jl:22-84): identical points on
    every rank (seed), per-rank pose shard (seed, rank)."""
    from tests.helpers import random_rotations
    dt = np.float32 if cfg["dtype"] == "f32" else np.float64
    rng_pts = np.random.Generator(np.random.PCG64(seed))
    points = np.asfortranarray((0.4 * rng_pts.standard_normal((cfg["n_in"], cfg["P"]))).astype(dt))
    w = rng_pts.random(cfg["P"])
    rng = np.random.Generator(np.random.PCG64([seed, 7919 + rank]))
    if cfg["n_in"] == 2:
        ang = rng.uniform(0, 2 * np.pi, cfg["B"])
        rot = np.empty((2, 2, cfg["B"]), dtype=dt, order="F")
        rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    else:
        rot = random_rotations(rng, cfg["n_in"], cfg["n_out"], cfg["B"], dt)
    tr = np.asfortranarray((0.1 * rng.standard_normal((cfg["n_out"], cfg["B"]))).astype(dt))
    d = dict(points=points, rotation=rot, translation=tr, background=None, out_weight=None, point_weight=None)
    if cfg["weights"]:
        d["background"] = np.arange(1, cfg["B"] + 1, dtype=dt)
        d["out_weight"] = (10 * rng.random(cfg["B"])).astype(dt)
        d["point_weight"] = (w / w.sum()).astype(dt)
    return d


def algorithmic_bytes(cfg):
    """Compulsory HBM bytes per pass: every array crosses HBM once."""
    s = 4 if cfg["dtype"] == "f32" else 8
    G = int(np.prod(cfg["grid"]))
    pts = (cfg["n_in"] + (1 if cfg["weights"] else 0)) * cfg["P"] * s
    pose = (cfg["n_out"] * cfg["n_in"] + cfg["n_out"] + 2) * cfg["B"] * s
    fwd = pts + pose + G * cfg["B"] * s
    bwd = G * cfg["B"] * s + 2 * (pts + pose)
    return fwd, bwd


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first NVML queries are slow: make them before the timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else {}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=0)
        return dict(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))


def host_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_time(cfg, inputs, n_poses, threads, steps=1, ds_dout=None):
    """Times the oracle port (the reference's CPU algorithm restated in C + OpenMP, same parallel structure:
    src/raster_pullback.jl:115-146) on the first n_poses poses.  Returns seconds per step (fwd, bwd)."""
    from oracle import oracle
    dt = np.float32 if cfg["dtype"] == "f32" else np.float64
    sl = lambda a: None if a is None else np.asfortranarray(a[..., :n_poses])
    args = (inputs["points"], sl(inputs["rotation"]), sl(inputs["translation"]), sl(inputs["background"]),
            sl(inputs["out_weight"]), inputs["point_weight"])
    if ds_dout is None:
        ds_dout = np.asfortranarray(np.random.default_rng(5).standard_normal(tuple(cfg["grid"]) + (n_poses,)).astype(dt))
    tf = tb = 0.0
    for _ in range(steps):
        if "fwd" in cfg["ops"]:
            t0 = time.perf_counter()
            oracle.raster(cfg["grid"], *args, dtype=dt, n_threads=threads)
            tf += time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle.raster_pullback(ds_dout, *args, dtype=dt, n_slabs=min(threads, n_poses))
        tb += time.perf_counter() - t0
    return tf / steps, tb / steps


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path.  Julia is not installed (here or on the
    GPU box), so the arm runs the oracle port of src/raster.jl / src/raster_pullback.jl with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    inputs = synth_inputs(cfg, 1000 + int(args.config[-1]), 0)
    n_poses = min(cfg["B"], max(threads, args.ref_poses))
    dt = np.float32 if cfg["dtype"] == "f32" else np.float64
    ds = np.asfortranarray(np.random.default_rng(5).standard_normal(tuple(cfg["grid"]) + (n_poses,)).astype(dt))
    for _ in range(args.warmup):
        cpu_port_time(cfg, inputs, n_poses, threads, 1, ds)
    t0 = time.perf_counter()
    tf, tb = cpu_port_time(cfg, inputs, n_poses, threads, args.steps, ds)
    wall = (time.perf_counter() - t0) / args.steps
    value = cfg["P"] * n_poses / (tf + tb)
    sample = f"{n_poses} of {cfg['B']} poses per step (poses are independent), all {cfg['P']} points"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * wall, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=cfg["dtype"], data="synthetic",
                config=dict(workload=f"{args.config}: {cfg['label']}", ops=cfg["ops"], sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample,
                                  fwd_splats_per_s=(cfg["P"] * n_poses / tf) if tf else None,
                                  bwd_splats_per_s=cfg["P"] * n_poses / tb,
                                  note="Julia absent: C/OpenMP restatement of the reference's CPU algorithm (oracle/)"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


class Workload:
    """One configuration resident on one GPU: inputs, outputs, the step (forward + pullback), its measurements."""

    def __init__(self, name, cfg, rank, dev, comm=None):
        import torch
        import dpr_b200
        from dpr_b200 import sharded
        self.name, self.cfg, self.dev = name, cfg, dev
        self.td = torch.float32 if cfg["dtype"] == "f32" else torch.float64
        seed = 1000 + int(name[-1]) + (100 if name.startswith("readme") else 0)
        self.inputs = synth_inputs(cfg, seed, rank)
        f = lambda a: None if a is None else dpr_b200.fortran(torch.from_numpy(np.ascontiguousarray(a)).to(dev))
        i = self.inputs
        self.points = f(i["points"])
        self.background, self.out_weight, self.point_weight = f(i["background"]), f(i["out_weight"]), f(i["point_weight"])
        # Two pose sets, alternated step by step: like an optimiser loop, no step sees the poses of the step before, so
        # nothing a step computes (the 3-d path's point bins, DPR_OPT_BINNING_CACHE) can be left over from the previous
        # one - only the pullback of a step may reuse what the forward of the SAME step binned (the rrule's call order).
        other = synth_inputs(cfg, seed + 50_000, rank)
        self.poses = [(f(i["rotation"]), f(i["translation"])), (f(other["rotation"]), f(other["translation"]))]
        grid, B = tuple(cfg["grid"]), cfg["B"]
        gen = torch.Generator(device=dev).manual_seed(seed * 31 + rank)
        self.ds_dout = dpr_b200.empty_f(grid + (B,), self.td, dev)
        self.ds_dout.normal_(generator=gen)
        self.do_fwd = "fwd" in cfg["ops"]
        self.out = dpr_b200.empty_f(grid + (B,), self.td, dev) if self.do_fwd else None
        self.drv = sharded.PoseShardedRaster(comm=comm)
        self.n_step = 0
        self.last = None

    def step(self):
        import dpr_b200
        rot, tr = self.poses[self.n_step & 1]
        self.n_step += 1
        if self.do_fwd:
            dpr_b200.raster_(self.out, self.points, rot, tr, self.background, self.out_weight, self.point_weight)
        self.last, _ = self.drv.raster_pullback_(self.ds_dout, self.points, rot, tr, self.background, self.out_weight,
                                                 self.point_weight)
        return self.last

    def measure(self, steps, warmup, sync_all, world, profiled_steps=None):
        """Times `steps` steps with CUDA events (no per-kernel instrumentation inside the timed region), then a second,
        shorter pass with the library's per-launch events for the kernel breakdown."""
        import torch
        import torch.distributed as dist
        import dpr_b200
        from dpr_b200 import _lib
        cfg = self.cfg
        for _ in range(warmup):
            self.step()
        sync_all()
        launches0 = dpr_b200.kernel_launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(steps):
            self.step()
        ev[1].record()
        sync_all()
        local_ms = ev[0].elapsed_time(ev[1]) / steps
        launches = dpr_b200.kernel_launch_count() - launches0
        paths = (dpr_b200.last_path(0) if self.do_fwd else "none", dpr_b200.last_path(1))
        ms_per_step, per_rank = local_ms, None
        if world > 1:
            t = torch.tensor([local_ms], dtype=torch.float64, device=self.dev)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [float(x.item()) for x in allt]
            ms_per_step = max(per_rank)
        # ---- kernel breakdown: a separate pass, every launch bracketed by events on its stream -------------------
        n_prof = min(steps, profiled_steps or 10)
        _lib.profile_enable(True)
        for _ in range(n_prof):
            self.step()
        sync_all()
        records = _lib.profile_records()
        _lib.profile_enable(False)
        kernels = {}
        for name, ms in records:
            kernels.setdefault(name, []).append(ms)
        kernel_ms = {k: sum(v) / n_prof for k, v in kernels.items()}              # per step (all launches of the kernel)
        kernel_launch_ms = {k: sum(v) / len(v) for k, v in kernels.items()}       # per launch
        launches_per_step = {k: len(v) / n_prof for k, v in kernels.items()}
        fwd_bytes, bwd_bytes = algorithmic_bytes(cfg)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        fwd_names = [k for k in kernel_ms if k.startswith("fwd_")]
        bwd_names = [k for k in kernel_ms if k.startswith("pullback_")]
        cand = {}
        if fwd_names and self.do_fwd:
            cand[max(fwd_names, key=kernel_ms.get)] = fwd_bytes
        if bwd_names:
            cand[max(bwd_names, key=kernel_ms.get)] = bwd_bytes
        dom = max(cand, key=lambda k: kernel_ms[k])
        bytes_per_launch = cand[dom] / launches_per_step[dom]
        achieved = bytes_per_launch / (kernel_launch_ms[dom] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(self.name, {}).get(dom)
        step_bytes = (fwd_bytes if self.do_fwd else 0) + bwd_bytes
        roofline = dict(bound="hbm", kernel=dom, achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                        algorithmic_bytes_per_launch=bytes_per_launch, kernel_ms=kernel_launch_ms[dom],
                        kernel_ms_median=sorted(kernels[dom])[len(kernels[dom]) // 2],
                        launches_per_step=launches_per_step[dom], peak_source=peak_src,
                        whole_step=dict(algorithmic_bytes=step_bytes, achieved=step_bytes / (ms_per_step * 1e-3) / 1e9,
                                        frac=step_bytes / (ms_per_step * 1e-3) / 1e9 / peak))
        return dict(ms_per_step=ms_per_step, per_rank_ms=per_rank, kernels_ms=kernel_ms, launches=launches, paths=paths,
                    roofline=roofline, splats_per_s=cfg["P"] * cfg["B"] * world / (ms_per_step * 1e-3),
                    fwd_splats_per_s=(cfg["P"] * cfg["B"] * world / (sum(v for k, v in kernel_ms.items() if k.startswith("fwd_") or k == "fill_background") * 1e-3))
                    if self.do_fwd and fwd_names else None)

    def free(self):
        import torch
        self.__dict__.clear()
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configurations (other_configs block)")
    ap.add_argument("--cpu-poses", type=int, default=2048, help="poses in the cpu_baseline sample")
    ap.add_argument("--ref-poses", type=int, default=512, help="poses per step of the --impl reference arm")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    claim_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    import dpr_b200
    from dpr_b200 import _lib, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # one process per GPU: run on the CPUs next to this GPU so the pinned host buffers of the end-to-end leg are
        # first-touched on the local NUMA node (8 ranks otherwise fight over one socket's memory and inter-socket links)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev)
    comm, comm_kind = None, "none"
    if world > 1:
        try:
            comm = sharded.DprComm(dev)
            comm_kind = ("dpr_comm_allreduce_sum: one-shot peer-memory kernels of libdpr.so (CUDA IPC over NVLink; NCCL only for the bootstrap)"
                         if comm.uses_peer_memory else "dpr_comm_allreduce_sum (NCCL via libdpr.so)")
        except Exception as e:   # NCCL could not be resolved inside the library: torch.distributed does the all-reduce
            comm, comm_kind = None, f"torch.distributed all_reduce ({type(e).__name__})"

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    wl = Workload(args.config, cfg, rank, dev, comm)
    P, B, grid = cfg["P"], cfg["B"], tuple(cfg["grid"])
    td = wl.td
    do_fwd = wl.do_fwd
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = wl.measure(args.steps, args.warmup, sync_all, world)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    ms_per_step = head["ms_per_step"]
    splats = P * B * world
    value = splats / (ms_per_step * 1e-3)

    # ---- multi-GPU: cost of the collective and a check of what it produced (outside the timed region) ----------------
    multi = None
    if world > 1:
        n_in = cfg["n_in"]
        packed = wl.drv.packed_buffer(n_in, P, td, dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 20
        sync_all()
        ev[0].record()
        for _ in range(reps):
            if comm is not None:
                comm.all_reduce_(packed)
            else:
                dist.all_reduce(packed)
        ev[1].record()
        sync_all()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_ms = float(t.item())
        # the all-reduced [d_points; d_point_weight] must equal the sum of the ranks' local pullbacks
        rot, tr = wl.poses[0]
        local = dpr_b200.raster_pullback_(wl.ds_dout, wl.points, rot, tr, wl.background, wl.out_weight, wl.point_weight)
        local_packed = torch.cat([local.points.t().reshape(-1), local.point_weight.reshape(-1)])
        gathered = [torch.empty_like(local_packed) for _ in range(world)]
        dist.all_gather(gathered, local_packed)
        want = torch.stack(gathered).to(torch.float64).sum(0)
        wl.n_step = 0
        wl.step()                                   # poses[0] again, through the sharded driver (with its all-reduce)
        got = wl.drv.packed_buffer(n_in, P, td, dev).to(torch.float64)
        err = float(((got - want).norm() / want.norm().clamp_min(1e-300)).item())
        tol = 1e-5 if cfg["dtype"] == "f32" else 1e-10
        multi = dict(allreduce_ms=allreduce_ms, allreduce_bytes=packed.numel() * packed.element_size(),
                     step_ms_min=min(head["per_rank_ms"]), step_ms_max=max(head["per_rank_ms"]), per_rank_step_ms=head["per_rank_ms"],
                     allreduce_check=dict(rel_l2_vs_sum_of_rank_partials=err, tol=tol, ok=bool(err <= tol)))

    # ---- end to end through the host-buffer C ABI (H2D of inputs and D2H of results inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        lib = _lib.load()
        suf = cfg["dtype"]
        points, (rotation, translation) = wl.points, wl.poses[0]
        background, out_weight, point_weight, ds_dout = wl.background, wl.out_weight, wl.point_weight, wl.ds_dout
        pin = lambda t: None if t is None else t.detach().cpu().pin_memory()
        h = dict(points=pin(points.t().contiguous()), rotation=pin(rotation.permute(2, 1, 0).contiguous()),
                 translation=pin(translation.t().contiguous()), background=pin(background), out_weight=pin(out_weight),
                 point_weight=pin(point_weight), ds_dout=pin(ds_dout.permute(*range(len(grid), -1, -1)).contiguous()))
        h_out = torch.empty(ds_dout.numel(), dtype=td).pin_memory() if do_fwd else None
        n_in, n_out = cfg["n_in"], cfg["n_out"]
        g = dict(dp=torch.empty(P * n_in, dtype=td).pin_memory(), drot=torch.empty(B * n_in * n_out, dtype=td).pin_memory(),
                 dtr=torch.empty(B * n_out, dtype=td).pin_memory(), dbg=torch.empty(B, dtype=td).pin_memory(),
                 dow=torch.empty(B, dtype=td).pin_memory(), dpw=torch.empty(P, dtype=td).pin_memory())
        import ctypes
        garr = (ctypes.c_int64 * len(grid))(*grid)
        p = lambda t: None if t is None else t.data_ptr()
        packed_dev = torch.empty((n_in + 1) * P, dtype=td, device=dev) if world > 1 else None

        def e2e_step():
            if do_fwd:
                _lib.check(getattr(lib, f"dpr_raster_forward_host_{suf}")(
                    n_in, n_out, garr, P, B, p(h["points"]), p(h["rotation"]), p(h["translation"]), p(h["background"]),
                    p(h["out_weight"]), p(h["point_weight"]), p(h_out)))
            _lib.check(getattr(lib, f"dpr_raster_pullback_host_{suf}")(
                n_in, n_out, garr, P, B, p(h["ds_dout"]), p(h["points"]), p(h["rotation"]), p(h["translation"]),
                p(h["out_weight"]), p(h["point_weight"]), p(g["dp"]), p(g["drot"]), p(g["dtr"]), p(g["dbg"]), p(g["dow"]),
                p(g["dpw"])))
            if world > 1:   # pose-sum across ranks: host -> device -> the library's collective -> host
                packed_dev[: n_in * P].copy_(g["dp"], non_blocking=True)
                packed_dev[n_in * P:].copy_(g["dpw"], non_blocking=True)
                if comm is not None:
                    comm.all_reduce_(packed_dev)
                else:
                    dist.all_reduce(packed_dev)
                g["dp"].copy_(packed_dev[: n_in * P], non_blocking=True)
                g["dpw"].copy_(packed_dev[n_in * P:], non_blocking=True)
                torch.cuda.synchronize(dev)

        def e2e_step_pipelined():
            """The same two calls, the forward non-blocking: ds_dout does not depend on `out` here, so the D2H copies of
            `out` and the H2D copies of ds_dout share the two directions of the host link."""
            ticket = ctypes.c_void_p()
            if do_fwd:
                _lib.check(getattr(lib, f"dpr_raster_forward_host_async_{suf}")(
                    n_in, n_out, garr, P, B, p(h["points"]), p(h["rotation"]), p(h["translation"]), p(h["background"]),
                    p(h["out_weight"]), p(h["point_weight"]), p(h_out), ctypes.byref(ticket)))
            rc_pb = getattr(lib, f"dpr_raster_pullback_host_{suf}")(
                n_in, n_out, garr, P, B, p(h["ds_dout"]), p(h["points"]), p(h["rotation"]), p(h["translation"]),
                p(h["out_weight"]), p(h["point_weight"]), p(g["dp"]), p(g["drot"]), p(g["dtr"]), p(g["dbg"]), p(g["dow"]),
                p(g["dpw"]))
            if do_fwd:
                _lib.check(lib.dpr_host_wait(ticket))
            _lib.check(rc_pb)
            if world > 1:
                packed_dev[: n_in * P].copy_(g["dp"], non_blocking=True)
                packed_dev[n_in * P:].copy_(g["dpw"], non_blocking=True)
                if comm is not None:
                    comm.all_reduce_(packed_dev)
                else:
                    dist.all_reduce(packed_dev)
                g["dp"].copy_(packed_dev[: n_in * P], non_blocking=True)
                g["dpw"].copy_(packed_dev[n_in * P:], non_blocking=True)
                torch.cuda.synchronize(dev)

        def time_e2e(fn):
            fn()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            sync_all()
            sec = (time.perf_counter() - t0) / args.e2e_steps
            if world > 1:
                t = torch.tensor([sec], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return sec

        dependent_s = time_e2e(e2e_step)
        e2e_s = time_e2e(e2e_step_pipelined) if do_fwd else dependent_s
        nbytes = lambda t: 0 if t is None else t.numel() * t.element_size()
        in_small = sum(nbytes(h[k]) for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight"))
        h2d = (in_small if do_fwd else 0) + in_small + nbytes(h["ds_dout"])
        d2h = nbytes(h_out) + sum(nbytes(v) for v in g.values())
        e2e = dict(value=splats / e2e_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=1e3 * e2e_s,
                   steps=args.e2e_steps,
                   api="dpr_raster_forward_host_async_* + dpr_raster_pullback_host_* + dpr_host_wait (pinned host buffers; the two "
                       "independent calls overlap, so both directions of the host link are busy)",
                   link_gbs=(h2d + d2h) / e2e_s / 1e9,
                   dependent=dict(value=splats / dependent_s, ms_per_step=1e3 * dependent_s,
                                  api="dpr_raster_forward_host_* then dpr_raster_pullback_host_* (blocking, back to back)"))
        lib.dpr_host_release()
        del h, h_out, g

    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu:
        threads = host_threads()
        n_poses = min(B, args.cpu_poses)
        cpu_port_time(cfg, wl.inputs, min(n_poses, threads), threads)  # warm the threads
        tf, tb = min((cpu_port_time(cfg, wl.inputs, n_poses, threads) for _ in range(2)), key=sum)   # best of two passes
        cpu_baseline = dict(value=P * n_poses / (tf + tb), unit=UNIT, cores=threads, kind="port",
                            sample=f"first {n_poses} of {B} poses, all {P} points, one pass (fwd {tf:.2f}s + bwd {tb:.2f}s)",
                            fwd_splats_per_s=(P * n_poses / tf) if tf else None, bwd_splats_per_s=P * n_poses / tb)

    # ---- every other BASELINE configuration, ten steps each (device-resident), so the driver's record carries them ------
    others = None
    if not args.no_others and args.config == "cfg2":
        wl.free()
        others = {}
        # one GPU: all of them; several ranks: config 5, the one BASELINE.json shards over 8 GPUs (2048 poses per rank)
        names = ("cfg1", "cfg3", "cfg4", "cfg5") if world == 1 else ("cfg5",)
        for name in names:
            try:
                w2 = Workload(name, CONFIGS[name], rank, dev, comm)
                r = w2.measure(10, 3, sync_all, world)
                others[name] = dict(workload=CONFIGS[name]["label"], ops=CONFIGS[name]["ops"], ms_per_step=r["ms_per_step"],
                                    splats_per_s=r["splats_per_s"], kernels_ms=r["kernels_ms"],
                                    roofline_frac=r["roofline"]["frac"], roofline_kernel=r["roofline"]["kernel"],
                                    whole_step_frac=r["roofline"]["whole_step"]["frac"], forward_path=r["paths"][0],
                                    pullback_path=r["paths"][1], per_rank_step_ms=r["per_rank_ms"])
                w2.free()
            except Exception as e:      # one configuration failing must not lose the headline line
                others[name] = dict(error=f"{type(e).__name__}: {e}"[:300])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=cfg["dtype"],
                    data="synthetic",
                    config=dict(workload=f"{args.config}: {cfg['label']}", ops=cfg["ops"], poses_per_gpu=B,
                                parallelism=f"pose-sharded x{world}, points replicated, 1 all-reduce of d_points+d_point_weight",
                                collective=comm_kind,
                                l2="inputs larger than L2 (out and ds_dout are 1.07 GB each per step; no flush needed)",
                                poses="two pose sets alternate step by step (no step repeats the previous step's inputs)",
                                timing="headline: CUDA events around the steps, no per-kernel instrumentation; kernels_ms: a separate pass",
                                forward_path=head["paths"][0], pullback_path=head["paths"][1]),
                    kernels_ms=head["kernels_ms"], fwd_splats_per_s=head["fwd_splats_per_s"],
                    roofline=head["roofline"], cpu_baseline=cpu_baseline, e2e=e2e, gpu_launches=head["launches"],
                    clocks=sampler.result(), multi_gpu=multi, other_configs=others)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
