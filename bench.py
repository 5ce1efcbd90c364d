#!/usr/bin/env python
"""bench.py -- cloud-march throughput (Mpix/s, ms/frame) of the B200 path, with its roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--filter exact|hw|hybrid]
    python bench.py --impl reference ...      # the CPU arm: the oracle port on all host cores

A "step" is one full-resolution frame (one MM_FULL dispatch of the cloud pass: every pixel marched).
Default workload: BASELINE.json configs[1] = C2, 1920x1080, midday sun, shipped CloudPlacement (mean
coverage 0.52), shipped noise volumes; see tests/scenes.py.  Inputs (4 textures, ~10 MB as bytes, 43 MB
resident incl. the float copies) are far smaller than L2 by design of the workload, so the L2 is flushed
between timed frames by writing a 256 MB buffer (config.l2: "flushed"); each frame is timed with its own
CUDA event pair on the launching stream and the flush is outside the pairs.

N > 1 (torchrun, one process per GPU): STRONG scaling -- the same frame is sharded row-cyclically
(row block 8 = the height of a thread block's pixel tile) over the ranks; every rank's kernel stores its pixels straight into rank 0's image over
NVLink (CUDA-IPC mapped peer memory), so there is no separate gather collective.  Time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TEXPEAK_QUADS_PER_CLK_PER_SM = 4
N_SM = 148


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--filter", default="hw", choices=["exact", "hw", "hybrid"],
                    help="hw (default): texture-unit filtering, parity against the oracle's bit-exact texture-unit model; "
                         "exact / hybrid: FP32 software filtering of the march samples, parity against the oracle's binary32 sampler")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, choices=[0, 1, 2, 4, 8], help="lanes sharing one ray: 0 = chosen per dispatch (default), 1, 2, 4, 8 (scheduling only)")
    ap.add_argument("--row-block", type=int, default=8, help="rows per row-cyclic shard block on N > 1 GPUs (8 = the block height of the march kernels; 4 measured 20 %% slower)")
    ap.add_argument("--animation", type=int, default=0, help="frame-parallel wind animation of N frames (BASELINE config 5): frame k on rank k %% world")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            pass

    def start(self):
        time.sleep(0.25)     # let the first samples land before the timed region starts

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.count(",") >= 5]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi gave no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = [name for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"])
                   if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(sm)}


def oracle_filter(ob, kernel_filter):
    """The oracle sampler that defines parity for a kernel filter mode."""
    return ob.OM_FILTER_TEXUNIT if kernel_filter == "hw" else ob.OM_FILTER_FP32


def workload_Q(oracle_binding, sc, rows_step, kernel_filter):
    """Algorithmic work per pixel: Q = N2D + 2*N3D bilinear-quad ops, from the oracle's counters on
    a row subsample of the same frame (every rows_step-th row)."""
    S = oracle_binding.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle_filter(oracle_binding, kernel_filter))
    t0 = time.time()
    _, cnt = S.march(sc["W"], sc["H"], row_begin=0, row_stride=rows_step, row_block=1)
    dt = time.time() - t0
    rows = cnt[::rows_step]
    npx = rows.shape[0] * rows.shape[1]
    return {"Q": float(rows[..., 1].mean() + 2.0 * rows[..., 2].mean()), "trips": float(rows[..., 0].mean()),
            "lit": float(rows[..., 3].mean()), "pixels": npx, "seconds": dt}


def _ref_shader_worker(job):
    """One process of the reference-shader timing: the reference's own compute-clouds.comp (oracle/_ref/libref_cc.so) over a
    slice of rows of a 1920x1080 frame, every pixel of those rows (4 phase calls per row group)."""
    import ctypes as C
    import _pkg
    import oracle_binding as ob
    import scenes
    cfg, rows, filt = job
    mm = _pkg.load_package()
    assets = scenes.load_assets()
    sc = scenes.make_scene(mm, cfg, assets)
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filt, pow_mode=ob.OM_POW_LIBM)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_cc.so"))
    ref.ref_cc_run.argtypes = [C.c_void_p] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    class Ctx(C.Structure):
        _fields_ = [("scene", C.c_void_p), ("filter", C.c_int)]
    ctx = Ctx(S.s, filt)
    out = np.zeros((1080, 1920, 4), np.float32)
    cam, sky = np.ascontiguousarray(sc["cam"], np.float32), np.ascontiguousarray(sc["sky"], np.float32)
    cb = C.cast(ob.lib().om_sample_callback, C.c_void_p)
    t0 = time.perf_counter()
    npx = 0
    for y in rows:
        for ox in range(4):
            sun = np.ascontiguousarray(sc["sun"], np.float32).copy()
            sun[11] = float((y % 4) * 4 + ox)                         # sun.color.a selects the 1-of-16 pixel phase (CC:292)
            ids = np.stack([np.arange(480, dtype=np.uint32), np.full(480, y // 4, np.uint32)], 1).copy()
            ref.ref_cc_run(ob._p(cam), ob._p(sun), ob._p(sky), cb, C.byref(ctx), ob._p(ids), len(ids), ob._p(out), None, None)
            npx += 480
    return npx, time.perf_counter() - t0, float(out[rows[0], :8].sum())


def time_reference_shader_build(cfg, filt, rows_step, cores):
    """Throughput of the reference's own shader text compiled for the CPU (single-threaded by construction: its uniform blocks are
    globals), one process per core; None when it is not available (no _ref build, or a frame the shader's hard-coded 1920x1080
    does not cover)."""
    import multiprocessing as mp
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cc.so")):
        return None
    try:
        rows = list(range(0, 1080, rows_step))
        chunks = [rows[i::cores] for i in range(cores) if rows[i::cores]]
        with mp.get_context("spawn").Pool(len(chunks)) as pool:
            t0 = time.perf_counter()
            res = pool.map(_ref_shader_worker, [(cfg, c, filt) for c in chunks])
            wall = time.perf_counter() - t0
        busy = max(r[1] for r in res)                      # slowest worker's compute time (excludes process start-up and asset loading)
        npx = sum(r[0] for r in res)
        return {"value": npx / busy / 1e6, "unit": "Mpix/s", "processes": len(chunks), "pixels": npx, "seconds_slowest_worker": busy,
                "seconds_wall_incl_startup": wall,
                "what": "the reference's own compute-clouds.comp, rewritten lexically to C++ and compiled for the CPU (oracle/_ref/libref_cc.so), one process per core"}
    except Exception as e:                                 # never let the extra measurement break the arm
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """CPU arm: the reference's GLSL cannot be built here (no glslang / Vulkan / lavapipe), so this times the
    oracle PORT of compute-clouds.comp on all host cores, on a bounded row sample of the same frame."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import _pkg
    import oracle_binding as ob
    import scenes
    mm = _pkg.load_package()
    assets = scenes.load_assets()
    sc = scenes.make_scene(mm, args.config, assets)
    cores = os.cpu_count()
    rows_step = max(1, int(sc["W"] * sc["H"] / 150e3))     # ~150k pixels per step
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle_filter(ob, args.filter))
    times = []
    npx = len(range(0, sc["H"], rows_step)) * sc["W"]
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        S.march(sc["W"], sc["H"], row_begin=0, row_stride=rows_step, row_block=1, counters=False)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    v = npx / dt / 1e6
    sample = f"every {rows_step}th row of the {sc['W']}x{sc['H']} {args.config} frame ({npx} px per step), OpenMP over rows"
    ref_build = None
    if sc["W"] == 1920 and sc["H"] == 1080:
        ref_build = time_reference_shader_build(args.config, oracle_filter(ob, args.filter), max(rows_step * 4, 1), cores)
    print(json.dumps({
        "impl": "reference", "metric": "cloud-march throughput", "value": v, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "ms_per_full_frame_extrapolated": sc["W"] * sc["H"] / v / 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} {sc['W']}x{sc['H']} full-resolution cloud march (every pixel), shipped CloudPlacement/CurlNoiseFBM/128^3/32^3 textures",
                   "filter": "oracle, texture-unit model sampler" if args.filter == "hw" else "oracle, binary32 sampler", "l2": "n/a (CPU)", "parallelism": f"{cores} host threads"},
        "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle port of compute-clouds.comp (bit-identical to the reference's own shader text executed on the CPU, tests/test_reference_shader.py); stands in for the reference shader on lavapipe, which cannot run here"},
        "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_shader_build": ref_build,
        "gpu_launches": 0,
    }))


def run_animation(args, mm, cs, sc, multigpu, dist, rank, world, local):
    """BASELINE config 5 as a job: N whole frames of a wind animation (sky.wind.w = 8*k), frame k rendered by
    rank k % world into its slot of rank 0's frame ring over NVLink.  One step = the whole animation."""
    import torch
    W, H, F = sc["W"], sc["H"], args.animation
    ring = multigpu.FrameRing(cs, rank, world, dist if world > 1 else None)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mine = ring.frames_of(F)
    sky = sc["sky"].copy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def job():
        for k in mine:
            sky[11] = 8.0 * k
            cs.updateUniformBuffers(sc["cam"], None, sky, sc["sun"])
            cs.dispatch(mm.MM_FULL, stream=stream.cuda_stream)

    for _ in range(max(1, args.warmup // 3)):
        job()
    barrier()
    times = []
    for _ in range(max(1, args.steps // 10)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        job()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    ms = sum(times) / len(times)
    if rank == 0:
        print(json.dumps({"metric": "cloud-march throughput", "value": F * W * H / ms / 1e3, "unit": "Mpix/s", "n_gpus": world, "steps": len(times),
                          "warmup": max(1, args.warmup // 3), "ms_per_step": ms, "ms_per_frame": ms / F, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": len(times) * len(mine),
                          "config": {"workload": f"{args.config} {W}x{H} x {F}-frame wind animation (sky.wind.w = 8k), frame-parallel: frame k on rank k % {world}",
                                     "filter": args.filter, "l2": "frames are 531 MB at 8K (> L2); no flush", "parallelism": f"frame-parallel x{world}, frames stored into rank 0's ring over NVLink"}}))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    ring.close()
    cs.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import _pkg
    import scenes
    mm = _pkg.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cloud pass has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    assets = scenes.load_assets()
    sc = scenes.make_scene(mm, args.config, assets)
    W, H = sc["W"], sc["H"]
    fmode = {"exact": mm.MM_FILTER_EXACT, "hw": mm.MM_FILTER_HW, "hybrid": mm.MM_FILTER_HYBRID}[args.filter]

    cs = mm.ComputeShader(local, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.setFilterMode(fmode)
    cs.setLanesPerRay(args.lanes)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])

    from project_marshmallow_b200 import multigpu
    if args.animation > 0:
        return run_animation(args, mm, cs, sc, multigpu, dist, rank, world, local)
    # output image lives on rank 0; other ranks map it through CUDA IPC and store into it over NVLink
    shared = multigpu.SharedFrame(cs, rank, world, dist if world > 1 else None)

    stream = torch.cuda.Stream()            # every launch and every event of the timed region is on this stream
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, Wm = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def frame_once(ev0=None, ev1=None):
        flush.fill_(1)                      # L2 flush between frames (outside the event pair)
        if ev0 is not None:
            ev0.record(stream)
        cs.dispatch(mm.MM_FULL, rank, world, args.row_block if world > 1 else 1, stream=stream.cuda_stream)
        if ev1 is not None:
            ev1.record(stream)

    for _ in range(Wm):
        frame_once()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        if world > 1:
            barrier()                       # all ranks start each sharded frame together
        frame_once(*evs[i])
    barrier()
    clocks = sampler.summary() if sampler else None
    per = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)      # frame time = slowest rank
    ms = float(per.mean())
    mpix = W * H / ms / 1e3

    # ---- e2e: the call a user makes with HOST buffers (uniforms in, image out), pinned host memory
    e2e = None
    if world > 1:
        # one page-locked shared-memory frame mapped by every rank: each kernel stores its pixels to rank 0's device image over
        # NVLink AND to this host frame over its own PCIe link, so the D2H transfer is fused into the march on all N GPUs
        hostframe = multigpu.SharedHostFrame(cs, rank, world, dist)
        n_e2e = max(3, K // 2)

        def e2e_once():
            cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
            cs.dispatch(mm.MM_FULL, rank, world, args.row_block, stream=stream.cuda_stream)
            barrier()                                   # every rank's stream has drained: the frame is complete in host memory
        for _ in range(2):
            e2e_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_once()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        checksum = float(hostframe.array[::97, ::89].sum()) if rank == 0 else 0.0      # the host frame is read by the CPU
        hostframe.close()
        e2e = {"value": W * H / dt / 1e6, "unit": "Mpix/s", "ms_per_frame": dt * 1e3, "h2d_bytes_per_step": 328 * world,
               "d2h_bytes_per_step": W * H * 16, "host_frame_checksum": checksum,
               "note": "uniform blocks from host on every rank; sharded march storing into rank 0's device image over NVLink and into one page-locked "
                       "shared host frame over each GPU's PCIe link (device->host transfer fused into the kernels); barrier; wall clock"}
    if world == 1:
        host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
        hnp = host.numpy()
        for _ in range(2):
            cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(3, K // 2)
        for _ in range(n_e2e):
            cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": W * H / dt / 1e6, "unit": "Mpix/s", "ms_per_frame": dt * 1e3, "h2d_bytes_per_step": 328,
               "d2h_bytes_per_step": W * H * 16, "note": "mm_render_to_host: uniform blocks from host, RGBA32F frame back to pinned host memory"}

    # ---- the reference's own per-frame cadence (VulkanApplication.cpp:1053-1071): reproject the previous image, then ONE
    # compute-clouds dispatch = 1/16 of the pixels (phase k % 16), ping-pong.  Reported beside the headline, which marches
    # EVERY pixel every frame (16 reference dispatches' worth of rays).
    cadence = None
    if world == 1:
        ping = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        pong = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        cs.bindOutput(ping.data_ptr())
        cs.dispatch(mm.MM_FULL, stream=stream.cuda_stream)
        cur, prev = pong, ping
        cev, k_rep, k_march = [], [], []
        for i in range(Wm + K):
            sun_i = sc["sun"].copy()
            sun_i[11] = float(i % 16)                                   # sun.color.a carries the pixel phase (CC:292)
            cs.updateUniformBuffers(sc["cam"], sc["cam"], sc["sky"], sun_i)
            cs.bindOutput(cur.data_ptr())
            cs.bindPrevious(prev.data_ptr())
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            cs.dispatchReproject(stream=stream.cuda_stream)
            cs.dispatch(mm.MM_PHASE16, stream=stream.cuda_stream)
            e1.record(stream)
            if i >= Wm:
                cev.append((e0, e1))
            cur, prev = prev, cur
        torch.cuda.synchronize()
        for i in range(K):                                              # the two kernels alone (no host launch gap between them)
            cs.bindOutput(cur.data_ptr())
            cs.bindPrevious(prev.data_ptr())
            flush.fill_(1)
            cs.dispatchReproject(stream=stream.cuda_stream)
            k_rep.append(cs.lastKernelMs())
            cs.dispatch(mm.MM_PHASE16, stream=stream.cuda_stream)
            k_march.append(cs.lastKernelMs())
            cur, prev = prev, cur
        cms = sum(a.elapsed_time(b) for a, b in cev) / len(cev)
        kr, km = sum(k_rep) / K, sum(k_march) / K
        cadence = {"ms_per_frame": kr + km, "frames_per_s": 1e3 / (kr + km), "reproject_kernel_ms": kr, "phase16_march_kernel_ms": km,
                   "stream_ms_per_frame_incl_host_launch_gaps": cms, "launches_per_frame": 2,
                   "note": "reference cadence: reprojection of the previous image + one MM_PHASE16 dispatch (1/16 of the pixels marched), ping-pong images, "
                           "L2 flushed between frames; ms_per_frame = sum of the two kernels' CUDA-event times, stream_ms = one event pair around both "
                           "launches issued from this Python binding"}

    # ---- texture-pipe ceiling measured live: L1-resident filtered fetches, bilinear (curl noise, 64 KB) and trilinear
    # (hi-res volume, 128 KB); reported beside the nominal 148 SM x 4 quads/clk x f_SM
    tex_peak = None
    if world == 1:
        try:
            ms2, q2 = cs.measureTexPeak(mm.MM_TEX_CURL, 4096)
            ms3, q3 = cs.measureTexPeak(mm.MM_TEX_HIRES, 4096)
            tex_peak = {"bilinear_2d_Gquad_s": q2 / 1e9, "trilinear_3d_Gquad_s": q3 / 1e9, "ms": [ms2, ms3],
                        "how": "mm_measure_tex_peak: 148 x 8 blocks x 256 threads x 4096 filtered fetches from an L1-resident texture, CUDA events"}
        except Exception as e:                      # the ceiling is supplementary: never let it break the bench line
            tex_peak = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peaks, which = measured_peaks()
        import oracle_binding as ob
        out = {
            "metric": "cloud-march throughput", "value": mpix, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms, "ms_per_frame": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} {W}x{H} full-resolution cloud march (every pixel), shipped CloudPlacement/CurlNoiseFBM/128^3/32^3 textures",
                       "filter": args.filter, "lanes_per_ray": args.lanes or "per dispatch", "l2": "flushed (256 MB write between frames)", "parallelism": f"row-cyclic x{world}, row block {args.row_block}" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": K, "e2e": e2e,
        }
        if cadence:
            out["reference_cadence"] = cadence
        if not args.no_cpu_baseline:
            rows_step = max(1, int(W * H / 300e3))
            wq = workload_Q(ob, sc, rows_step, args.filter)
            texpeak = N_SM * TEXPEAK_QUADS_PER_CLK_PER_SM * peaks["sm_max_mhz"] * 1e6
            achieved = wq["Q"] * W * H / (ms * 1e-3)
            out["roofline"] = {"bound": "tex", "achieved": achieved / 1e9, "peak": texpeak / 1e9, "unit": "Gquad/s", "frac": achieved / texpeak,
                               "traffic": None, "Q_quads_per_pixel": wq["Q"], "loop_trips_per_pixel": wq["trips"], "lit_steps_per_pixel": wq["lit"],
                               "peak_source": f"148 SM x 4 bilinear quads/clk x sm_max_mhz ({which} clock); algorithmic quads from the oracle's fetch counters",
                               "hbm_floor_ms": W * H * 16 / (peaks["hbm_gbs"] * 1e9) * 1e3}
            if tex_peak:
                out["roofline"]["measured_tex_peak"] = tex_peak
                best = max(tex_peak.get("bilinear_2d_Gquad_s", 0.0), tex_peak.get("trilinear_3d_Gquad_s", 0.0))
                if best > 0:
                    out["roofline"]["frac_of_measured_peak"] = achieved / 1e9 / best
            # second roofline: the march is FP32 / instruction-issue bound.  Warp-instructions per frame of this
            # exact command come from the committed ncu capture (profiles/); peak = 148 SM x 4 schedulers x f_SM.
            prof = os.path.join(ROOT, "profiles", f"r01b_{args.filter}.summary.csv")        # capture of the current kernel revision
            if not os.path.exists(prof):
                prof = os.path.join(ROOT, "profiles", f"r01_final_{args.filter}.summary.csv")
            if os.path.exists(prof) and args.config == "C2" and world == 1:
                rows = {l.split(",")[0]: l.strip().split(",") for l in open(prof) if l.count(",") >= 2}
                unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                try:        # DRAM bytes of one launch of this command, from the committed ncu --set full capture
                    out["roofline"]["traffic"] = sum(float(rows[k][2]) * unit[rows[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    out["roofline"]["traffic_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, " + os.path.relpath(prof, ROOT) + "); the 33 MB image mostly stays in L2 at kernel end"
                except (KeyError, ValueError):
                    pass
                winst = [float(l.split(",")[2]) for l in open(prof) if l.startswith("smsp__inst_executed.sum,")]
                if winst:
                    peak_issue = N_SM * 4 * peaks["sm_max_mhz"] * 1e6
                    out["roofline_issue"] = {"bound": "instruction issue (FP32)", "achieved": winst[0] / (ms * 1e-3) / 1e9, "peak": peak_issue / 1e9,
                                             "unit": "Gwarp-inst/s", "frac": winst[0] / (ms * 1e-3) / peak_issue,
                                             "warp_instructions_per_frame": winst[0], "source": os.path.relpath(prof, ROOT)}
            v = wq["pixels"] / wq["seconds"] / 1e6
            out["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"every {rows_step}th row of the same frame ({wq['pixels']} px), oracle port with counters, OpenMP",
                                   "note": "the port is bit-identical to the reference's own shader text executed on the CPU (tests/test_reference_shader.py)"}
        print(json.dumps(out))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    shared.close()
    cs.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
