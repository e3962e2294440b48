#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native WavefrontPathTracer pass.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Cornell box, 1920x1080, wavefront path tracing, max depth 10,
NEE + MIS, rr 0.8.  One STEP = one frame = `--spp` samples per pixel through beginFrame + render.
Metric: Mrays/s = (closest-hit rays + shadow rays popped from the queues) / device time
(SURVEY.md section 8d).  `value` is timed with the film in HBM; `e2e` goes through the C ABI with
HOST buffers (camera struct in, RGBA32F film out through pinned memory) inside the timed region.

Multi-GPU (N ranks, one per GPU; kiraray_b200/multigpu.py): the work is split by image tile and/or
by spp slice (`--partition spp|tile|hybrid`, default spp: rank r renders frame indices r+1, r+1+N,
... -- the reference accumulates spp across frames), the scene is replicated, and the films are
summed with ONE NCCL reduce to rank 0 per step, inside the timed region.  With the spp split every
GPU renders a full frame per step, so per-GPU work is fixed: scaling = "weak".

`--impl reference`: the reference's own KRR_CALLABLE integrator code compiled host-side
(oracle/_ref) driven by the CPU restatement of the wavefront stages, on all host cores, on a bounded
row-band sample of the same workload per step.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

# Workloads = BASELINE.json configs[1..4].  `geom` = geometry lower bound in bytes per ray (one
# compressed 80-B node per BLAS / TLAS level + one triangle), so bytes/ray = 508 (queue traffic in the reference's
# SoA layout) + geom: 752 (cbox), 1232 (20 M triangles), 1392 (10 k instances); the volumetric config adds 4 B per
# tentative collision (one density fetch) and reports collisions per ray.
# `batch` = frames in flight per step ("frame_batch": F consecutive frame indices rendered by the same launches, film =
# their mean): a stage launch lasts as long as its slowest ray, and F frames share that latency.  One STEP = one
# render() = F frames x spp samples per pixel.
WORKLOADS = {
    "cbox": dict(name="cbox_1080p_depth10_nee", config=1, size=(1920, 1080), max_depth=10, spp=8, geom=244, batch=4),
    "tess20m": dict(name="tess20m_disney_1k_emitters_1080p_depth10_nee", config=2, size=(1920, 1080), max_depth=10, spp=2, geom=724, batch=8),
    "smoke": dict(name="cbox_density_grid_in_mist_1080p_depth15_nee", config=3, size=(1920, 1080), max_depth=15, spp=1, geom=244, batch=4),
    "inst10k": dict(name="inst10k_two_level_srt_motionblur_refit_4k_depth5_nee", config=4, size=(3840, 2160), max_depth=5, spp=1, geom=884, batch=4),
}
RR = 0.8
QUEUE_BYTES_PER_RAY = 508.0


def stage_bytes(geom):
    """per-stage split of the algorithmic figure (DESIGN.md "Roofline accounting"), bytes per unit the stage processes"""
    return {"closest": 120 + 32 + 232 + geom, "scatter": 232 + 32 + 32 + 92 + 120, "shadow": 92 + 32 + geom, "medium": 316 + 84}


def load_ncu(workload, kernel):
    """Counters of `kernel` from the committed `ncu --set full` capture of this workload (profiles/traffic.json,
    written by tools/ncu_summary.py traffic): DRAM bytes per launch, issue-slot utilisation, warps active, SIMT
    efficiency.  {} when not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return {}
    d = json.load(open(p))
    return d.get(workload, {}).get(kernel, {}) or (d.get(kernel, {}) if workload == "cbox" else {})


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region through NVML in-process
    (the same fields as the nvidia-smi line of B200_PROFILING.md; forking nvidia-smi from a process
    that holds a CUDA context stalls kernel launches for milliseconds, which would distort the
    measurement it is supposed to qualify)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.err = index, [], False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML handle by UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = None
            for i in range(nv.nvmlDeviceGetCount()):
                hi = nv.nvmlDeviceGetHandleByIndex(i)
                u = nv.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, reasons))
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = max([r[1] for r in self.rows] or [0])
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = [n for n, b in bits.items() if any(r[2] & b for r in self.rows)]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


class Workload:
    """Scene, camera(s) and pass parameters of one BASELINE config.  Scene synthesis is input generation only."""

    def __init__(self, key, spp=None, scale=1.0, batch=None):
        import kiraray_b200 as krr
        from kiraray_b200 import scenes
        self.key, self.spec = key, WORKLOADS[key]
        self.batch = batch or self.spec["batch"]
        self.W, self.H = self.spec["size"]
        self.max_depth = self.spec["max_depth"]
        self.spp = spp or self.spec["spp"]
        self.geom = self.spec["geom"]
        self.bytes_per_ray = QUEUE_BYTES_PER_RAY + self.geom
        self.app = self.builder = None
        self.extra = {}
        if key in ("cbox", "smoke"):
            cfg = "cbox.json" if key == "cbox" else "config4_smoke.json"
            self.app = krr.HostApp(os.path.join(ROOT, "assets", "configs", cfg), asset_root=ROOT)
            self.app.set_resolution(self.W, self.H)
            self.app.set_wfpt_params(spp=self.spp, max_depth=self.max_depth, rr=RR, nee=True)
            self.desc, self.params = self.app.scene_desc(), dict(self.app.wfpt_params())
            self._cam = self.app.camera()
        elif key == "tess20m":
            self.builder = scenes.tessellated_scene(n_objects=max(8, int(200 * scale)), tris_per_object=max(2000, int(100_000 * scale)), n_emissive=1000)
            self.desc, self.params = self.builder.build(), dict(spp=self.spp, max_depth=self.max_depth, rr=RR, nee=True)
            self._cam = scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), self.W / self.H)
            self.extra = {"triangles": self.builder.triangle_count()}
        elif key == "inst10k":
            ng = max(4, int(round(100 * scale ** 0.5)))
            self.builder, _ = scenes.instanced_scene(n_blas=16, tris_per_blas=max(500, int(20_000 * scale)), n_groups=ng, per_group=ng, motion=True, time=0.5)
            self.desc, self.params = self.builder.build(), dict(spp=self.spp, max_depth=self.max_depth, rr=RR, nee=True)
            # every frame has its own shutter interval (the animation advances), so begin_frame re-fits the TLAS
            # boxes of all moving instances inside the timed region: "per-frame BVH refit"
            self._cams = [scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), self.W / self.H, shutter_open=0.3 + 0.05 * k, shutter_time=0.05) for k in range(8)]
            self._cam = self._cams[0]
            self.extra = {"instances": ng * ng, "triangles": self.builder.triangle_count()}

    def camera(self, step=0):
        return self._cams[step % len(self._cams)] if self.key == "inst10k" else self._cam

    def config(self, spp):
        return dict({"workload": self.spec["name"], "baseline_config": self.spec["config"], "width": self.W, "height": self.H,
                     "max_depth": self.max_depth, "rr": RR, "nee": True, "spp_per_frame": spp, "frames_per_step": self.batch,
                     "spp_per_step": spp * self.batch}, **self.extra)


def cpu_reference_rate(wl, spp, rows, threads=0, frames=1):
    """Times the CPU oracle on a band of `rows` image rows of the workload, `frames` frames.  Returns (Mrays/s, info)."""
    import oracle_binding as ob
    kind = "reference"
    orc = ob.Oracle(wl.desc, kind)
    W, H = wl.W, wl.H
    rows = min(rows, H)
    r0 = (H - rows) // 2
    rays, seconds = 0, 0.0
    for f in range(frames):
        # use_bvh=2: static instances through the oracle's BVH, moving instances through a per-frame BVH over their boxes for the
        # frame's shutter interval (oracle/driver.cpp buildMovingTlas) -- what a CPU tracer would do; use_bvh=1 visits every
        # moving instance per ray (that mode exists to verify the kernels' motion bounds, tests/test_gpu_motion.py)
        res = orc.render(wl.camera(f), W, H, frame_index=1 + f, spp=spp, max_depth=wl.max_depth, rr=RR, use_bvh=2, threads=threads, rows=(r0, r0 + rows))
        rays += res["stats"]["closest_rays"] + res["stats"]["shadow_rays"]
        seconds += res["seconds"]
    orc.close()
    return rays / seconds / 1e6, {"kind": kind, "rays": rays, "seconds": seconds,
                                  "sample": f"rows {r0}..{r0 + rows} of {H} ({rows * W} pixels), {spp} spp, {frames} frame{'s' if frames > 1 else ''}"}


_STDOUT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    stdout when NCCL_DEBUG is set in the environment), so everything else is sent to stderr and the line is
    written to the saved descriptor at the end."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, (json.dumps(line) + "\n").encode())


L2_NOTE = "inputs larger than L2: the per-step queue + pixel-state working set (0.9 GB at 1080p, 3.6 GB at 4K) exceeds the 126 MB L2; no flush between steps"


def full_config(wl, args, world):
    """The `config` object of BOTH arms (identical by construction: the driver compares them)."""
    from kiraray_b200.multigpu import make_partition
    part = make_partition(0, max(world, 1), wl.H, args.partition)
    return dict(wl.config(args.spp or wl.spp), parallelism=part.describe() + (", fixed total spp (strong)" if args.strong else ""), l2=L2_NOTE)


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = Workload(args.workload, args.spp, args.scene_scale, args.frame_batch)
    spp = wl.spp
    cores = os.cpu_count() or 1
    # ~1/8 of the cbox frame per step: a few seconds of CPU work; the tree scenes cost ~10x more per ray
    rows = args.ref_rows or {"cbox": 135, "tess20m": 32, "smoke": 24, "inst10k": 270}[args.workload]
    vals = []
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference_rate(wl, spp, rows)
        if i >= args.warmup:
            vals.append((v, info["seconds"], info["rays"]))
    rays = sum(v[2] for v in vals)
    secs = sum(v[1] for v in vals)
    value = rays / secs / 1e6
    line = {"impl": "reference", "metric": "Mrays/s (primary+shadow+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": full_config(wl, args, world),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cbox", choices=sorted(WORKLOADS), help="BASELINE.json config: cbox = configs[1] (the headline), tess20m = [2], smoke = [3], inst10k = [4]")
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per frame (one step = one frame); 0 = the workload's default")
    ap.add_argument("--frame-batch", type=int, default=0, help="frames in flight per step (pass parameter frame_batch); 0 = the workload's default")
    ap.add_argument("--scene-scale", type=float, default=1.0, help="< 1 shrinks the triangle / instance counts of the synthetic scenes (smoke tests only)")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows of the frame the CPU reference renders per step")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames the cpu_baseline leg renders")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=None, help="length of the sustained (steady-state clocks) measurement that follows the K timed steps; 0 = skip; default 5 s on one GPU, skipped on N > 1")
    ap.add_argument("--params", default="", help="extra pass parameters as JSON (tuning switches, e.g. '{\"pdl\": false}')")
    ap.add_argument("--partition", default="spp", choices=["spp", "tile", "hybrid"], help="multi-GPU work split (N > 1)")
    ap.add_argument("--reduce", default="abi", choices=["abi", "torch"], help="N > 1 film reduction: the library's own NCCL entry points (krr_wfpt_reduce_film; the product path) or torch.distributed.reduce (diagnostic)")
    ap.add_argument("--strong", action="store_true", help="fixed TOTAL work: the spp of a step are divided among the ranks' spp slices (scaling = strong)")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.min_seconds is None:
        args.min_seconds = 5.0 if world == 1 else 0.0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import kiraray_b200 as krr
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from kiraray_b200.multigpu import FilmReducer, make_partition
    wl = Workload(args.workload, args.spp, args.scene_scale, args.frame_batch)
    W, H = wl.W, wl.H
    part = make_partition(rank, world, H, args.partition)
    spp = wl.spp
    if args.strong:
        # fixed TOTAL work per step (spp x frames samples per pixel): every spp slice takes 1/S of it, as many frames
        # in flight as it can keep and fewer samples per frame
        total = wl.spp * wl.batch
        if total % part.spp_slices:
            raise SystemExit(f"--strong: {total} samples per pixel per step are not divisible by the {part.spp_slices} spp slices")
        share = total // part.spp_slices
        wl.batch = max(b for b in range(1, wl.batch + 1) if share % b == 0)
        spp = share // wl.batch
    # debug_taps off: the C ABI's default (the ctypes test binding turns the parity taps on by default)
    gpu = krr.Wfpt(params={**wl.params, "spp": spp, "debug_taps": False, "frame_batch": wl.batch, **(json.loads(args.params) if args.params else {})})
    t0 = time.time()
    gpu.set_scene(wl.desc)
    accel_build_s = time.time() - t0
    gpu.resize(W, H)
    if part.tiles > 1:
        gpu.set_partition(*part.rows)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    film = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    film_host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = film_host.numpy()
    # the one exchange step: film sum-reduce to rank 0 over NVLink, by the product's own NCCL entry point
    reducer = FilmReducer(gpu, part, dist, use_torch=args.reduce == "torch")

    def frame_of(step):
        return part.frame_index(step, batch=wl.batch)  # first of the F consecutive frame indices of this step

    def step_device(i):
        gpu.begin_frame(frame_of(i), wl.camera(i), sptr)
        gpu.render(film.data_ptr(), sptr)
        reducer.reduce(film, sptr)

    film_host2 = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_bufs = [host_np, film_host2.numpy()]

    def step_e2e_sync(i):
        gpu.begin_frame(frame_of(i), wl.camera(i), sptr)
        gpu.render_to_host(host_np, sptr)        # render + D2H of the film + stream sync, every step

    def step_e2e(i):
        gpu.begin_frame(frame_of(i), wl.camera(i), sptr)  # camera struct: host -> device (kernel arguments)
        if dist is None:
            # render + D2H of the film into one of two pinned host buffers on the copy stream: the read-back of
            # step i overlaps the rendering of step i + 1 (krr_wfpt_render_to_host_async); the timed region ends
            # with krr_wfpt_wait_host, i.e. when the film of every step is in host memory
            gpu.render_to_host_async(host_bufs[i & 1], sptr)
        else:
            # N > 1: render, NCCL reduce into one of two reduced films, and (rank 0) the pipelined read-back of
            # that film on the copy stream while the next step renders
            reducer.render_reduce_to_host_async(host_bufs[i & 1], sptr)

    def drain_e2e():
        if dist is None:
            gpu.wait_host()
        else:
            reducer.wait_host()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, drain=None, steps=None, first=None):
        steps = steps or args.steps
        first = args.warmup if first is None else first
        for i in range(args.warmup):
            fn(i)
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(first + (i % args.steps))
        if drain:
            drain()  # host-side wait for the copy stream; e1 is recorded after the last film has landed
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident timing ----
    ms = timed(step_device)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches_per_step = gpu.stats()["kernel_launches"]
    # ray counts: deterministic per frame index, gathered with an untimed replay of the timed frames
    rays = 0
    collisions = 0
    for i in range(args.steps):
        gpu.begin_frame(frame_of(args.warmup + i), wl.camera(args.warmup + i), sptr)
        gpu.render(film.data_ptr(), sptr)
        st = gpu.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
        collisions += st.get("medium_collisions", 0)
    # ---- end-to-end timing (host buffers) ----
    ms_e2e = timed(step_e2e, drain_e2e)
    ms_e2e_sync = timed(step_e2e_sync) if dist is None else None
    # ---- sustained: the same K frames over and over for >= --min-seconds, clocks sampled over the whole run ----
    sustained = None
    if args.min_seconds > 0:
        # EVERY rank must run the same number of steps (each step ends in a collective): size the leg from the slowest
        # rank's time, not from the local one (tiles of a hybrid split cost different amounts; even equal work differs by
        # a fraction of a per cent, enough to round to another step count on one rank and hang the film reduce)
        ms_all = ms
        if dist is not None:
            tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_all = float(tm[0])
        n_sus = max(args.steps, int(args.min_seconds * 1e3 / max(ms_all / args.steps, 1e-3)) + 1)
        n_sus = (n_sus + args.steps - 1) // args.steps * args.steps  # whole cycles of the K counted frames
        sus_sampler = ClockSampler(local)
        sus_sampler.start()
        ms_sus = timed(step_device, steps=n_sus)
        sus_sampler.stop_flag = True
        sus_sampler.join(timeout=2)
        sustained = (ms_sus, n_sus, sus_sampler.summary())
    # ---- per-stage profile of one step (events around every launch; not part of `value`) ----
    gpu.set_profiling(True)
    gpu.begin_frame(frame_of(args.warmup), wl.camera(args.warmup), sptr)
    gpu.render(film.data_ptr(), sptr)
    torch.cuda.synchronize()
    stages = gpu.stage_times()
    gpu.set_profiling(False)
    st = gpu.stats()

    pixels = (part.rows[1] - part.rows[0]) * W
    t = torch.tensor([ms, ms_e2e, float(rays), float(pixels), sustained[0] if sustained else 0.0, float(collisions)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, rays, pixels, collisions = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3]), float(tsum[5])
        if sustained:
            sustained = (float(tmax[4]),) + sustained[1:]
    if rank == 0:
        peak, peak_src = load_peaks()
        SB = stage_bytes(wl.geom)
        bytes_per_ray = wl.bytes_per_ray + (4.0 * collisions / rays if collisions else 0.0)
        value = rays / (ms * 1e-3) / 1e6
        e2e = rays / (ms_e2e * 1e-3) / 1e6
        # dominant kernel = the stage with the largest share of the profiled step.  The fused trace launch
        # (k_trace_fused) traces the shadow rays of depth d and the closest rays of depth d + 1: its units are
        # rays of both kinds and its algorithmic bytes the sum of the two stages' figures.
        cands = [k for k in ("closest", "scatter", "shadow", "trace", "medium") if stages[k]["launches"]]
        dom = max(cands, key=lambda k: stages[k]["ms"])
        names = {"closest": "k_trace_closest", "scatter": "k_scatter<Disney>", "shadow": "k_trace_shadow_tr" if "medium" in cands else "k_trace_shadow",
                 "trace": "k_trace_fused", "medium": "k_medium_sample"}
        fused = stages["trace"]["launches"] > 0  # then k_trace_closest only runs depth 0

        def stage_units(k):
            if k == "trace":
                n_c, n_s = st["closest_rays"] - st["closest_by_depth"][0], st["shadow_rays"]
                return n_c + n_s, n_c * SB["closest"] + n_s * SB["shadow"]
            u = {"closest": st["closest_by_depth"][0] if fused else st["closest_rays"], "scatter": st["scatter_items"], "shadow": st["shadow_rays"],
                 "medium": st["medium_sample_items"]}[k]
            return u, u * SB[k]
        units, dom_bytes = stage_units(dom)
        total_ms = sum(v["ms"] for v in stages.values())
        dom_gbs = dom_bytes / (stages[dom]["ms"] * 1e-3) / 1e9
        # the same figure for every traced / shaded stage, each with the ncu counters of its committed capture:
        # measured DRAM bytes per launch and the issue-side counters that actually bound these kernels
        per_stage = {}
        for k in cands:
            u, b = stage_units(k)
            gbs = b / (stages[k]["ms"] * 1e-3) / 1e9
            ncu = load_ncu(args.workload, names[k].split("<")[0])
            per_stage[names[k]] = {"achieved": gbs, "frac": gbs / peak, "share": stages[k]["ms"] / total_ms if total_ms else 0,
                                   "units_per_step": u, "launches_per_step": stages[k]["launches"], "traffic": ncu.get("dram_bytes_per_launch"),
                                   "issue_active_pct": ncu.get("issue_active_pct"), "warps_active_pct": ncu.get("warps_active_pct"),
                                   "simt_efficiency": ncu.get("simt_efficiency"), "dram_pct_of_peak": ncu.get("dram_pct_of_peak"),
                                   "bound": ncu.get("bound", "issue/latency (see profiles/)")}
        dncu = load_ncu(args.workload, names[dom].split("<")[0])
        step_ncu = load_ncu(args.workload, "__step__")
        roofline = {"bound": "hbm", "binding_resource": dncu.get("bound", "issue/latency: DRAM traffic is far below the algorithmic bytes (queues live in L2); see issue_active_pct / simt_efficiency"),
                    "kernel": names[dom],
                    "achieved": dom_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom_gbs / peak,
                    "traffic": dncu.get("dram_bytes_per_launch"),
                    "issue_active_pct": dncu.get("issue_active_pct"), "warps_active_pct": dncu.get("warps_active_pct"),
                    "simt_efficiency": dncu.get("simt_efficiency"),
                    "dram_bytes_per_step": step_ncu.get("dram_bytes_per_step"),
                    "dram_bytes_per_ray": (step_ncu["dram_bytes_per_step"] / (rays / max(1, world) / args.steps)) if step_ncu.get("dram_bytes_per_step") and rays else None,
                    "kernel_ms_per_step_serialised": step_ncu.get("kernel_ms_per_step_serialised"),
                    "algorithmic_bytes_per_unit": dom_bytes / max(1, units), "units_per_step": units, "launches_per_step": stages[dom]["launches"],
                    "avg_launch_ms": stages[dom]["ms"] / max(1, stages[dom]["launches"]),
                    "stage_share": {k: (v["ms"] / total_ms if total_ms else 0) for k, v in stages.items()}, "per_stage": per_stage,
                    "note": "algorithmic bytes = the reference's SoA queue layout + geometry lower bound (SURVEY 8d); a stage that moves less than that layout (depth-0 ray items here hold origin + direction only) can exceed frac 1",
                    "bytes_per_ray": bytes_per_ray, "collisions_per_ray": (collisions / rays) if collisions else None,
                    "pipeline_achieved": value * 1e6 * bytes_per_ray / 1e9 / max(1, world), "pipeline_frac": value * 1e6 * bytes_per_ray / 1e9 / max(1, world) / peak}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rows = args.ref_rows or {"cbox": H, "tess20m": 96, "smoke": 64, "inst10k": 1080}[args.workload]
            frames = args.ref_frames or (6 if args.workload == "cbox" else 1)  # ~10-20 s of CPU work on 16 cores
            v, info = cpu_reference_rate(wl, spp, rows, frames=frames)
            cpu = {"value": v, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": info["kind"], "sample": info["sample"]}
        spp_s = spp * wl.batch * args.steps * (pixels / (W * H)) / (ms * 1e-3)  # full-frame samples per pixel per second, all ranks
        line = {"metric": "Mrays/s (primary+shadow+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": full_config(wl, args, world),
                "spp_per_s": spp_s, "accel_build_s": round(accel_build_s, 3), "bvh": {"nodes": st["bvh_nodes"], "triangles": st["bvh_triangles"], "tlas_nodes": st["tlas_nodes"]},
                "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(krr.KrrCameraData), "d2h_bytes_per_step": W * H * 16,
                        "readback": ("pipelined: film of step i copied to pinned host memory on a copy stream while step i + 1 renders; timed region ends when every film is on the host"
                                     if world == 1 else
                                     "film sum-reduced to rank 0 by torch.distributed.reduce, then (rank 0) copied to pinned host memory, per step" if reducer.use_torch else
                                     "film sum-reduced to rank 0 by ncclReduce on the render stream (krr_wfpt_reduce_film), then (rank 0) copied to pinned host memory on a copy stream while step i + 1 renders; timed region ends when every film is on the host"),
                        "film_reduce": None if world == 1 else ("torch.distributed: " + getattr(reducer, "fallback_reason", "requested")) if reducer.use_torch else "krr_wfpt_reduce_film (NCCL inside the library)",
                        "value_sync_per_step": (rays / (ms_e2e_sync * 1e-3) / 1e6) if ms_e2e_sync else None},
                "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary()}
        if sustained:
            ms_sus, n_sus, clk = sustained
            line["sustained"] = {"value": rays * (n_sus / args.steps) / (ms_sus * 1e-3) / 1e6, "unit": "Mrays/s", "seconds": ms_sus * 1e-3, "steps": n_sus,
                                 "spp": n_sus * spp * wl.batch, "clocks": clk}
        emit(line)
    reducer.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
