#!/usr/bin/env python
"""bench.py — cost-volume throughput (build + aggregate + WTA) in Mpix*disp/s.

One "step" = the full dense-matching job of the workload: every reference view of the synthetic
8-view refractive scene (BASELINE.json configs[3], "cfg4", the configuration north_star quotes
its target on) through sr_run_view: stage (1) refractive tap-volume build, stages (2)+(3)
support-weight aggregation with fused WTA.  Metric = sum over reference views of H*W*D / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg3|small]

N > 1 is launched by torchrun (one rank per GPU); reference views are partitioned across ranks
(north_star: "by reference view for multi-view runs"), no collective on the data path; the total
work is fixed, so scaling is "strong".  `--impl reference` times the CPU oracle (the reference
cannot be compiled here) on a bounded row band with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from stereoreconstruction_b200 import scenes, sharding, types as T  # noqa: E402

METRIC = "cost-volume Mpix*disp/s (build+aggregate+WTA)"
UNIT = "Mpix*disp/s"


def workload(name):
    """Returns dict(w,h,V,D,params,cams,max_nbrs,desc)."""
    if name == "cfg4":
        w, h, V, D = 1920, 1080, 8, 256
    elif name == "cfg5":
        w, h, V, D = 3840, 2160, 8, 512
    elif name == "small":
        w, h, V, D = 480, 270, 8, 64
    elif name == "cfg3":
        w, h, V, D = 1920, 1080, 2, 256
    else:
        raise SystemExit(f"unknown workload {name}")
    if name == "cfg3":
        cams, _ = scenes.rectified_pair(w, h, z0=100.0)
        P = T.default_params(False, 100.0, 500.0, D, radius=16, weight_kind=T.SR_WEIGHT_ADAPTIVE)
        desc = "cfg3: synthetic rectified 1920x1080 pair, 256 disparities, AdaptiveWeight 33x33, NCC, both directions"
        surf = scenes.HeightField(z0=167.0, amp=20.0, lx=25.0, ly=18.0)
        cell = 3.5 * 167.0 / cams[0].K[0]
        seed = 1234
    else:
        cams = scenes.arc_cameras(V, w, h)
        P = T.default_params(True, 350.0, 650.0, D)  # GeodesicWeight r=2, NCC, 3 neighbours (reference defaults)
        desc = (f"{name}: synthetic refractive {V}-view {w}x{h}, {D} depth hypotheses, tilted planar interface "
                f"(n=1.333), lens distortion, GeodesicWeight r=2, 3 nearest neighbour views, all reference views")
        surf = scenes.HeightField(z0=0.0, amp=25.0, lx=90.0, ly=70.0)
        cell = 3.5 * 500.0 / cams[0].K[0]
        seed = 4321 if name != "cfg5" else 8765
    return dict(name=name, w=w, h=h, V=V, D=D, params=P, cams=cams, desc=desc, surf=surf, cell=cell, seed=seed)


def neighbours_for(wl, ctx=None):
    if wl["name"] == "cfg3":
        return [[1], [0]]
    return scenes.nearest_neighbours(wl["cams"], 3)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([time.time()] + [c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self, t_begin=None, t_end=None):
        """Samples inside [t_begin, t_end] (the timed region); when the region is shorter than two
        sampling periods, all samples since the sampler started (warm-up steps + timed region, the
        same load) are used and `window` says so."""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows, window = self.rows, "warmup+timed"
        if t_begin is not None:
            inside = [r for r in self.rows if t_begin <= r[0] <= t_end]
            if len(inside) >= 2:
                rows, window = inside, "timed"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def cpu_sample(wl, imgs, rows, steps=1, warmup=0, target_s=12.0):
    """Times the CPU oracle (all host threads) on a row band of reference view 0 of the workload.
    rows <= 0: a short probe sizes the band so that one timed pass takes about `target_s`."""
    from oracle import oracle_api as O
    sc = O.Scene(wl["cams"], imgs)
    nb = neighbours_for(wl)
    P = T.SrParams.from_buffer_copy(wl["params"])

    last = {}

    def run(nrows):
        r0 = (wl["h"] - nrows) // 2
        P.row_begin, P.row_end = r0, r0 + nrows
        t0 = time.perf_counter()
        if wl["name"] == "cfg3":
            od, oi, ob, _ = sc.twoview_label(P, 0, 1)
        else:
            od, oi, ob, _, _ = sc.mvs_view(P, 0, nb[0])
        dt = time.perf_counter() - t0
        last.update(depth=od[r0:r0 + nrows], index=oi[r0:r0 + nrows], best=ob[r0:r0 + nrows], r0=r0, rows=nrows)
        return dt, r0

    if rows <= 0:
        probe = max(2, min(wl["h"], O.num_threads()))  # one row per thread
        tp, _ = run(probe)
        rows = int(max(probe, min(wl["h"], probe * target_s / max(tp, 1e-3))))
        rows -= rows % max(1, O.num_threads())  # whole rows per thread under the static schedule
        rows = max(rows, probe)
    times = []
    r0 = 0
    for i in range(warmup + steps):
        dt, r0 = run(rows)
        if i >= warmup:
            times.append(dt)
    units = rows * wl["w"] * wl["D"]
    t = sum(times) / len(times)
    cpu_sample.last = last  # the band's outputs: compared with the GPU's for the same rows (parity at full size)
    return units / t / 1e6, t, O.num_threads(), f"{rows} rows x {wl['w']} px x {wl['D']} labels of reference view 0 (rows {r0}..{r0 + rows})"


def reference_sample(wl, imgs, rows, steps=1, warmup=0, target_s=12.0):
    """Times the REFERENCE'S OWN MultiViewStereo::computeInitialEstimate (stereo/multiviewstereo.cpp compiled
    where it lies, oracle/_ref/libref.so; its tbb::parallel_for over rows on all host threads) on a row
    band of reference view 0 of the workload: the reference view's mask is reduced to the band (the class
    skips pixels outside its mask), the neighbours are the ones its runTask() rule selects.  This is the
    reference's live formulation — candidates are the pixels of the rasterised epipolar curve that the D
    depth labels span — so a unit of work is still one (pixel, depth label).  Returns None when the
    prebuilt library is not there (then the oracle port is timed instead)."""
    from oracle import oracle_api as O
    if wl["name"] == "cfg3" or O.ref_lib() is None:
        return None
    P = wl["params"]
    nb = neighbours_for(wl)
    h = wl["h"]
    threads = [1]

    def timed(nrows, r0):
        # a fresh task object per pass: narrowing view 0's mask to the band cannot be undone
        band = O.RefMVS(wl["cams"], imgs, None, P.min_depth, P.max_depth, P.num_levels, 5.0, image_scale=P.image_scale)
        threads[0] = band.num_threads()
        for v in range(wl["V"]):
            band.set_neighbours(v, nb[v])
        band.mask_rows(0, r0, r0 + nrows)
        t0 = time.perf_counter()
        band.initial_estimate(0)
        dt = time.perf_counter() - t0
        band.close()
        return dt

    if rows <= 0:
        import os as _os
        probe = max(2, min(h, _os.cpu_count() or 2))  # about one row per thread
        tp = timed(probe, (h - probe) // 2)
        rows = int(max(probe, min(h, probe * target_s / max(tp, 1e-3))))
        rows -= rows % max(1, threads[0])
        rows = max(rows, probe)
    r0 = (h - rows) // 2
    times = []
    for i in range(warmup + steps):
        dt = timed(rows, r0)
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    units = rows * wl["w"] * wl["D"]
    return units / t / 1e6, t, threads[0], (f"{rows} rows x {wl['w']} px x {wl['D']} depth levels of reference view 0 "
                                         f"(rows {r0}..{r0 + rows}), MultiViewStereo::computeInitialEstimate of the reference itself")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--views", type=int, default=0, help="only the first N reference views (profiling aid; 0 = all)")
    ap.add_argument("--partition", default="views", choices=["views", "rows"],
                    help="multi-GPU: deal reference views to ranks (default) or give every rank a row band of every view")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)
    w, h, V, D = wl["w"], wl["h"], wl["V"], wl["D"]
    cpu_rows = args.cpu_rows  # 0: sized by a probe to ~12 s of CPU work

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        imgs = scenes.noise_images(V, w, h, wl["seed"])  # content does not change the CPU work
        kind, note = "reference", ("the reference's own MultiViewStereo (stereo/multiviewstereo.cpp compiled where it lies, "
                                   "oracle/_ref; its tbb::parallel_for over rows); each step = the bounded sample")
        got = reference_sample(wl, imgs, cpu_rows, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        if got is None:
            kind, note = "port", "CPU oracle (restated reference, OpenMP over rows); each step = the bounded sample"
            got = cpu_sample(wl, imgs, cpu_rows, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        val, t, cores, sample = got
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "note": note},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    import torch.distributed as dist
    from stereoreconstruction_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Context(local_rank)
    stream = torch.cuda.Stream()  # an explicit stream: the library launches on it, the events time it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # synthetic, photo-consistent inputs rendered through the GPU's own unproject
    cams = wl["cams"]
    blank = [np.zeros((h, w, 4), np.uint8)] * V
    ctx.set_views(cams, blank, None)
    ctx.set_params(wl["params"])
    cache = os.environ.get("SR_BENCH_IMAGE_CACHE")  # A/B aid: several runs in one session render the scene once
    if cache and os.path.exists(cache):
        imgs = list(np.load(cache)["imgs"])
    else:
        imgs = scenes.render_views(V, lambda v: ctx.unproject_grid(v), wl["surf"], wl["seed"], wl["cell"])
        if cache:
            np.savez(cache, imgs=np.stack(imgs))
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    imgs_p = [p.numpy() for p in pinned]
    ctx.set_views(cams, imgs_p, None)
    nbrs = neighbours_for(wl)
    ref_views = list(range(V if args.views <= 0 else min(V, args.views)))
    if args.partition == "rows" and world > 1:  # every rank: rows [b0, b1) of every reference view
        my_views = list(ref_views)
        b0, b1 = sharding.row_bands(h, world)[rank]
        wl["params"].row_begin, wl["params"].row_end = b0, b1
        ctx.set_params(wl["params"])
    else:
        my_views = [v for v in sharding.partition_views(V, world)[rank] if v in ref_views]
    units_total = len(ref_views) * h * w * D

    def step():
        for v in my_views:
            ctx.run_view(v, nbrs[v])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launch_count()
    ctx.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.finish(t_begin, t_end)
    stages = ctx.stage_ms()
    if os.environ.get("SR_MATCH_STATS"):
        print("match stats:", ctx.match_stats(), file=sys.stderr)
        print("build stats:", ctx.build_stats(), file=sys.stderr)
    ctx.set_profiling(False)
    launches = ctx.launch_count() - launches0
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = units_total / (ms_per_step * 1e-3) / 1e6

    # end-to-end through the C ABI with host buffers: H2D of all images + run + D2H of the index maps
    idx_host = [torch.empty((h, w), dtype=torch.int32).pin_memory().numpy() for _ in my_views]

    # a rank uploads the views it computes and their neighbours; the others stay camera-only
    needed = set(my_views) | {n for v in my_views for n in nbrs[v]}
    imgs_e2e = [imgs_p[v] if v in needed else None for v in range(V)]

    def e2e_step():
        ctx.set_views(cams, imgs_e2e, None)
        ctx.set_params(wl["params"])
        for v in my_views:
            ctx.run_view(v, nbrs[v])
        for k, v in enumerate(my_views):
            ctx.depth_index(v, out=idx_host[k])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tmax = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_val = units_total / e2e_s / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (match_kernel: streams the tap volume once, fused WTA)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    n_nbr = len(nbrs[my_views[0]])
    n_match = max(1, stages["match_launches"])
    match_ms = stages["match_ms"] / n_match
    build_ms = stages["build_ms"] / n_match
    # 8 B per pixel*disparity (volume written once by build, read once by
    # aggregate+WTA) + H*W*(4*(1+N_nbr)+4) per reference view; the match launch owns the read half
    # and the per-pixel terms.
    alg_bytes_match = h * w * D * 4 + h * w * (4 * (1 + n_nbr) + 4)
    alg_bytes_view = h * w * D * 8 + h * w * (4 * (1 + n_nbr) + 4)
    achieved = alg_bytes_match / (match_ms * 1e-3) / 1e9
    pipe_gbs = alg_bytes_view / ((match_ms + build_ms) * 1e-3) / 1e9
    traffic = None
    kname = "match_mvs_screen_kernel" if wl["name"] != "cfg3" else "match_kernel"
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(wl["name"], {}).get(kname)
    except (OSError, ValueError):
        pass
    roofline = {
        "bound": "hbm", "kernel": kname + " (support-weight aggregation of the photo-consistency cost + fused WTA)",
        "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
        "peak_source": peak_src, "match_ms_per_view": match_ms, "build_ms_per_view": build_ms,
        "pipeline_achieved_gbs": pipe_gbs, "pipeline_frac": pipe_gbs / peak_gbs,
        "binding_bound": ("instruction issue / latency, not HBM: ~165 warp-instructions per (pixel, label, neighbour) in the "
                          "match kernel and ~210 in the refractive build (profiles/, DESIGN.md section 5); HBM sees "
                          "only the 4-byte tap per (pixel, label, neighbour)"),
    }

    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline leg runs on rank 0 at N = 1 only
        val, t, cores, sample = cpu_sample(wl, imgs, cpu_rows)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds": t}
        # The oracle's outputs for that band are the checker of the GPU's at the benchmark's full size.
        o = cpu_sample.last
        if 0 in my_views and o:
            ra, rb = o["r0"], o["r0"] + o["rows"]
            gi, gd, gb = ctx.depth_index(0)[ra:rb], ctx.depth(0)[ra:rb], ctx.best_cost(0)[ra:rb]
            mism = gi != o["index"]
            lab = ~mism & (o["index"] >= 0)
            with np.errstate(invalid="ignore"):
                rel = np.abs(gb[lab] - o["best"][lab]) / np.maximum(np.abs(o["best"][lab]), 1e-300)
            cpu["parity"] = {
                "pixels": int(mism.size), "index_mismatch_rate": float(mism.mean()),
                "depth_equal_where_index_equal": bool(((gd == o["depth"]) | (np.isnan(gd) & np.isnan(o["depth"])))[~mism].all()),
                "max_rel_cost_diff": float(rel.max()) if rel.size else None, "labelled_fraction": float(lab.mean()),
            }
        # ... and the reference's own MultiViewStereo (oracle/_ref, when the prebuilt library is there) on a band
        # of the same view: its live curve formulation, timed beside the port (optional: never fails the line)
        try:
            got = reference_sample(wl, imgs, 0, target_s=8.0)
        except Exception as e:  # noqa: BLE001
            got = None
            cpu["reference_itself"] = {"unavailable": str(e)[:200]}
        if got:
            cpu["reference_itself"] = {"value": got[0], "unit": UNIT, "cores": got[2], "kind": "reference",
                                       "sample": got[3], "seconds": got[1]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "views_per_rank": len(my_views), "neighbours": n_nbr,
                   "l2": "inputs larger than L2 (tap volume %.1f GB per view streams through HBM)" % (n_nbr * h * w * D * 4 / 1e9),
                   "partition": ("row bands of every reference view" if (args.partition == "rows" and world > 1)
                                 else "reference views round-robin over ranks"),
                   "precision": ("every output (index, depth, winning cost) is decided in FP64; FP32 only screens labels "
                                 "that provably cannot win, and labels between exactly projected anchors are interpolated "
                                 "under a pixel-boundary guard (DESIGN.md section 3)")},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(len(needed) * h * w * 4),
                "d2h_bytes_per_step": int(len(my_views) * h * w * 4), "seconds_per_step": e2e_s},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
