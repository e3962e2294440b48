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

WORKLOAD = "cbox_1080p_depth10_nee"
W, H = 1920, 1080
MAX_DEPTH, RR = 10, 0.8
# algorithmic bytes per ray for the Cornell box (reference SoA layout + geometry LB)
BYTES_PER_RAY = 752.0
# per-stage split of that figure (DESIGN.md "Roofline accounting"), bytes per unit the stage processes
STAGE_BYTES = {"closest": 120 + 32 + 232 + 244, "scatter": 232 + 32 + 32 + 92 + 120, "shadow": 92 + 32 + 244}


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic); None when not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")


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


def make_app(spp):
    import kiraray_b200 as krr
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    app.set_resolution(W, H)
    app.set_wfpt_params(spp=spp, max_depth=MAX_DEPTH, rr=RR, nee=True)
    return app


def cpu_reference_rate(app, spp, rows, threads=0, frames=1):
    """Times the CPU oracle on a band of `rows` image rows of the workload, `frames` frames.  Returns (Mrays/s, info)."""
    import oracle_binding as ob
    kind = "reference" if ob.available("reference") else "port"
    orc = ob.Oracle(app.scene_desc(), kind)
    cam = app.camera()
    r0 = (H - rows) // 2
    rays, seconds = 0, 0.0
    for f in range(frames):
        res = orc.render(cam, W, H, frame_index=1 + f, spp=spp, max_depth=MAX_DEPTH, rr=RR, use_bvh=True, threads=threads, rows=(r0, r0 + rows))
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    app = make_app(args.spp)
    cores = os.cpu_count() or 1
    rows = args.ref_rows or 135  # ~1/8 of the frame per step: a few seconds of CPU work
    vals = []
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference_rate(app, args.spp, rows)
        if i >= args.warmup:
            vals.append((v, info["seconds"], info["rays"]))
    rays = sum(v[2] for v in vals)
    secs = sum(v[1] for v in vals)
    value = rays / secs / 1e6
    line = {"impl": "reference", "metric": "Mrays/s (primary+shadow+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "max_depth": MAX_DEPTH, "rr": RR, "nee": True, "spp_per_step": args.spp},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--spp", type=int, default=8, help="samples per pixel per frame (one step = one frame)")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows of the frame the CPU reference renders per step")
    ap.add_argument("--ref-frames", type=int, default=6, help="frames the cpu_baseline leg renders (full frame each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--params", default="", help="extra pass parameters as JSON (tuning switches, e.g. '{\"pdl\": false}')")
    ap.add_argument("--partition", default="spp", choices=["spp", "tile", "hybrid"], help="multi-GPU work split (N > 1)")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
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

    from kiraray_b200.multigpu import make_partition, reduce_film
    part = make_partition(rank, world, H, args.partition)
    app = make_app(args.spp)
    cam = app.camera()
    # debug_taps off: the C ABI's default (the ctypes test binding turns the parity taps on by default)
    gpu = krr.Wfpt(params=dict(app.wfpt_params(), debug_taps=False, **(json.loads(args.params) if args.params else {})))
    gpu.set_scene(app.scene_desc())
    gpu.resize(W, H)
    if part.tiles > 1:
        gpu.set_partition(*part.rows)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    film = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    film_host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = film_host.numpy()

    def frame_of(step):
        return part.frame_index(step)

    def step_device(i):
        gpu.begin_frame(frame_of(i), cam, sptr)
        gpu.render(film.data_ptr(), sptr)
        reduce_film(film, part, dist)  # film accumulation over NVLink (the one exchange step)

    film_host2 = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_bufs = [host_np, film_host2.numpy()]

    def step_e2e_sync(i):
        gpu.begin_frame(frame_of(i), cam, sptr)
        gpu.render_to_host(host_np, sptr)        # render + D2H of the film + stream sync, every step

    def step_e2e(i):
        gpu.begin_frame(frame_of(i), cam, sptr)  # camera struct: host -> device (kernel arguments)
        if dist is None:
            # render + D2H of the film into one of two pinned host buffers on the copy stream: the read-back of
            # step i overlaps the rendering of step i + 1 (krr_wfpt_render_to_host_async); the timed region ends
            # with krr_wfpt_wait_host, i.e. when the film of every step is in host memory
            gpu.render_to_host_async(host_bufs[i & 1], sptr)
        else:
            gpu.render(film.data_ptr(), sptr)
            reduce_film(film, part, dist)
            if rank == 0:
                film_host.copy_(film, non_blocking=True)
            stream.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, drain=None):
        for i in range(args.warmup):
            fn(i)
        if drain:
            drain()
        barrier()
        rays = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            fn(args.warmup + i)
        if drain:
            drain()  # host-side wait for the copy stream; e1 is recorded after the last film has landed
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        return ms

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident timing ----
    # rays of the timed steps: the counters are per frame, so re-render the same frames untimed below
    ms = timed(step_device)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches_per_step = gpu.stats()["kernel_launches"]
    # ray counts: deterministic per frame index, gathered with an untimed replay of the timed frames
    rays = 0
    for i in range(args.steps):
        gpu.begin_frame(frame_of(args.warmup + i), cam, sptr)
        gpu.render(film.data_ptr(), sptr)
        st = gpu.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
    # ---- end-to-end timing (host buffers) ----
    ms_e2e = timed(step_e2e, gpu.wait_host if dist is None else None)
    ms_e2e_sync = timed(step_e2e_sync) if dist is None else None
    # ---- per-stage profile of one step (events around every launch; not part of `value`) ----
    gpu.set_profiling(True)
    gpu.begin_frame(frame_of(args.warmup), cam, sptr)
    gpu.render(film.data_ptr(), sptr)
    torch.cuda.synchronize()
    stages = gpu.stage_times()
    gpu.set_profiling(False)
    st = gpu.stats()

    pixels = (part.rows[1] - part.rows[0]) * W
    t = torch.tensor([ms, ms_e2e, float(rays), float(pixels)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, rays, pixels = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3])
    if rank == 0:
        peak, peak_src = load_peaks()
        value = rays / (ms * 1e-3) / 1e6
        e2e = rays / (ms_e2e * 1e-3) / 1e6
        # dominant kernel = the stage with the largest share of the profiled step.  The fused trace launch
        # (k_trace_fused) traces the shadow rays of depth d and the closest rays of depth d + 1: its units are
        # rays of both kinds and its algorithmic bytes the sum of the two stages' figures.
        cands = [k for k in ("closest", "scatter", "shadow", "trace") if stages[k]["launches"]]
        dom = max(cands, key=lambda k: stages[k]["ms"])
        names = {"closest": "k_trace_closest", "scatter": "k_scatter<Disney>", "shadow": "k_trace_shadow", "trace": "k_trace_fused"}
        fused = stages["trace"]["launches"] > 0  # then k_trace_closest only runs depth 0

        def stage_units(k):
            if k == "trace":
                n_c, n_s = st["closest_rays"] - st["closest_by_depth"][0], st["shadow_rays"]
                return n_c + n_s, n_c * STAGE_BYTES["closest"] + n_s * STAGE_BYTES["shadow"]
            u = {"closest": st["closest_by_depth"][0] if fused else st["closest_rays"], "scatter": st["scatter_items"], "shadow": st["shadow_rays"]}[k]
            return u, u * STAGE_BYTES[k]
        units, dom_bytes = stage_units(dom)
        total_ms = sum(v["ms"] for v in stages.values())
        dom_gbs = dom_bytes / (stages[dom]["ms"] * 1e-3) / 1e9
        # the same figure for every traced / shaded stage (the two big ones are within a few per cent of each
        # other, so which one is "dominant" can change from run to run)
        per_stage = {}
        for k in cands:
            u, b = stage_units(k)
            gbs = b / (stages[k]["ms"] * 1e-3) / 1e9
            per_stage[names[k]] = {"achieved": gbs, "frac": gbs / peak, "share": stages[k]["ms"] / total_ms if total_ms else 0,
                                   "units_per_step": u, "launches_per_step": stages[k]["launches"], "traffic": load_traffic(names[k].split("<")[0])}
        roofline = {"bound": "hbm", "kernel": names[dom],
                    "achieved": dom_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom_gbs / peak,
                    "traffic": load_traffic(names[dom].split("<")[0]),
                    "algorithmic_bytes_per_unit": dom_bytes / max(1, units), "units_per_step": units, "launches_per_step": stages[dom]["launches"],
                    "avg_launch_ms": stages[dom]["ms"] / max(1, stages[dom]["launches"]),
                    "stage_share": {k: (v["ms"] / total_ms if total_ms else 0) for k, v in stages.items()}, "per_stage": per_stage,
                    "note": "algorithmic bytes = the reference's SoA queue layout + geometry lower bound (SURVEY 8d); a stage that moves less than that layout (depth-0 ray items here hold origin + direction only) can exceed frac 1",
                    "pipeline_achieved": value * 1e6 * BYTES_PER_RAY / 1e9 / max(1, world), "pipeline_frac": value * 1e6 * BYTES_PER_RAY / 1e9 / max(1, world) / peak}
        cpu = None
        if not args.no_cpu_baseline:
            v, info = cpu_reference_rate(app, args.spp, args.ref_rows or H, frames=args.ref_frames)  # ~10-15 s of CPU work on 16 cores
            cpu = {"value": v, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": info["kind"], "sample": info["sample"]}
        spp_s = args.spp * args.steps * (pixels / (W * H)) / (ms * 1e-3)  # full-frame samples per pixel per second, all ranks
        line = {"metric": "Mrays/s (primary+shadow+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "width": W, "height": H, "max_depth": MAX_DEPTH, "rr": RR, "nee": True,
                           "spp_per_step": args.spp, "parallelism": part.describe(),
                           "l2": "per-step queue + pixel-state working set (~0.9 GB) exceeds the 126 MB L2", "spp_per_s": spp_s},
                "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(krr.KrrCameraData), "d2h_bytes_per_step": W * H * 16,
                        "readback": "pipelined: film of step i copied to pinned host memory on a copy stream while step i + 1 renders; timed region ends when every film is on the host" if world == 1 else "after the NCCL film reduce, rank 0, stream sync per step",
                        "value_sync_per_step": (rays / (ms_e2e_sync * 1e-3) / 1e6) if ms_e2e_sync else None},
                "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary()}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
