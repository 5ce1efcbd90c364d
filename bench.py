#!/usr/bin/env python
"""bench.py -- cloud-march throughput (Mpix/s, ms/frame) of the B200 path, with its roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3] [--filter hw|hybrid|exact]
    python bench.py --impl reference ...      # the CPU arm: the oracle port on all host cores

A "step" is one full-resolution frame (one MM_FULL dispatch of the cloud pass: every pixel marched).
Default workload: BASELINE.json configs[2] = C3, 3840x2160, low sunset sun, shipped CloudPlacement and noise volumes
(configs/C3.json, tests/scenes.py) -- the configuration BASELINE's targets (4K < 1 ms, >= 85 % at 8 GPUs) are quoted on; C2
(configs[1], 1920x1080 noon) is reported beside it in `extra` at N = 1.  The textures (~10 MB) are far smaller than L2 by
design of the workload, so the L2 is flushed between timed frames by writing a 256 MB buffer (config.l2: "flushed"); each
frame is timed with its own CUDA event pair on the launching stream and the flush is outside the pairs.

N > 1 (torchrun, one process per GPU): STRONG scaling -- the same frame is sharded row-cyclically (row block 8) over the
ranks; every rank's kernel stores its pixels straight into rank 0's image over NVLink (CUDA-IPC mapped peer memory), so
there is no separate gather collective.  Time = max over ranks.  Rank 0 re-renders the frame alone and the bench ASSERTS that
the sharded frame equals it bit for bit; `frame_sha256` is printed at every N (equal hashes = identical frames).
"""
import argparse
import hashlib
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
    ap.add_argument("--config", default="C3")
    ap.add_argument("--filter", default="hw", choices=["exact", "hw", "hybrid"],
                    help="hw (default): texture-unit filtering, parity against the oracle's bit-exact texture-unit model; "
                         "exact / hybrid: FP32 software filtering of the march samples, parity against the oracle's binary32 sampler")
    ap.add_argument("--arith", default="ieee", choices=["ieee", "fma"],
                    help="arithmetic definition: ieee (default) = one rounding per operator; fma = the lexical contraction rule (its own oracle)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the N = 1 side blocks (parity_grade, sampler tail, C2, cadence, tex peak, sustained)")
    ap.add_argument("--lanes", type=int, default=0, choices=[0, 1, 2, 4, 8], help="lanes sharing one ray: 0 = chosen per dispatch (default), 1, 2, 4, 8 (scheduling only)")
    ap.add_argument("--scheduler", default="auto", choices=["auto", "static", "persistent"], help="static grid (K1) or persistent warps + dynamic tile queue (K1p); scheduling only")
    ap.add_argument("--refill", type=int, default=0, choices=[0, 8, 16, 32], help="K1p: dead lanes that trigger a refill (0/32 = a tile at a time)")
    ap.add_argument("--row-block", type=int, default=8, help="rows per row-cyclic shard block on N > 1 GPUs")
    ap.add_argument("--animation", type=int, default=0, help="frame-parallel wind animation of N frames (BASELINE config 5): frame k on rank k %% world")
    ap.add_argument("--sustained-seconds", type=float, default=2.5)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

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
        out = {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(sm)}
        try:
            out["power_w_max"] = max(float(r[6]) for r in rows if len(r) > 6)
        except ValueError:
            pass
        return out


def oracle_filter(ob, kernel_filter):
    """The oracle sampler that defines parity for a kernel filter mode."""
    return ob.OM_FILTER_TEXUNIT if kernel_filter == "hw" else ob.OM_FILTER_FP32


def host_threads():
    """Threads the CPU arms use: every host core.  Passed to the oracle EXPLICITLY -- torchrun exports OMP_NUM_THREADS=1."""
    return os.cpu_count() or 1


def workload_Q(oracle_binding, sc, rows_step, kernel_filter, arith="ieee"):
    """Algorithmic work per pixel: Q = N2D + 2*N3D bilinear-quad ops, from the oracle's counters on
    a row subsample of the same frame (every rows_step-th row)."""
    S = oracle_binding.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle_filter(oracle_binding, kernel_filter),
                             arith=oracle_binding.OM_ARITH_FMA if arith == "fma" else oracle_binding.OM_ARITH_IEEE)
    t0 = time.time()
    _, cnt = S.march(sc["W"], sc["H"], row_begin=0, row_stride=rows_step, row_block=1, nthreads=host_threads())
    dt = time.time() - t0
    rows = cnt[::rows_step]
    npx = rows.shape[0] * rows.shape[1]
    return {"Q": float(rows[..., 1].mean() + 2.0 * rows[..., 2].mean()), "trips": float(rows[..., 0].mean()),
            "lit": float(rows[..., 3].mean()), "pixels": npx, "seconds": dt}


def _ref_shader_worker(job):
    """One process of the reference-shader timing: the reference's own compute-clouds.comp (oracle/_ref/libref_cc.so) over a
    slice of rows of a 1920x1080 frame, every pixel of those rows (4 phase calls per row group)."""
    import ctypes as C
    import oracle_binding as ob
    import scenes
    cfg, rows, filt = job
    sc = scenes.scene_from_config(cfg, scenes.load_assets())
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
    oracle PORT of compute-clouds.comp on all host cores, on a bounded row sample of the same frame.  Inputs come from
    configs/<name>.json and the asset fixtures: nothing of the product library is loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_binding as ob
    import scenes
    sc = scenes.scene_from_config(args.config, scenes.load_assets())
    cores = host_threads()
    rows_step = max(1, int(sc["W"] * sc["H"] / 150e3))     # ~150k pixels per step
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle_filter(ob, args.filter),
                 arith=ob.OM_ARITH_FMA if args.arith == "fma" else ob.OM_ARITH_IEEE)
    times = []
    npx = len(range(0, sc["H"], rows_step)) * sc["W"]
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        S.march(sc["W"], sc["H"], row_begin=0, row_stride=rows_step, row_block=1, counters=False, nthreads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    v = npx / dt / 1e6
    sample = f"every {rows_step}th row of the {sc['W']}x{sc['H']} {args.config} frame ({npx} px per step), OpenMP over rows, {cores} threads requested explicitly"
    ref_build = time_reference_shader_build("C2", oracle_filter(ob, args.filter), 28, cores)       # the shader hard-codes 1920x1080: timed on C2
    print(json.dumps({
        "impl": "reference", "metric": "cloud-march throughput", "value": v, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "ms_per_full_frame_extrapolated": sc["W"] * sc["H"] / v / 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} {sc['W']}x{sc['H']} full-resolution cloud march (every pixel), shipped CloudPlacement/CurlNoiseFBM/128^3/32^3 textures",
                   "filter": "oracle, texture-unit model sampler" if args.filter == "hw" else "oracle, binary32 sampler", "l2": "n/a (CPU)",
                   "parallelism": f"{cores} host threads (omp num_threads passed explicitly; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')} is ignored)"},
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
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
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
    clocks = sampler.summary() if sampler else None
    ms = sum(times) / len(times)
    if rank == 0:
        print(json.dumps({"metric": "cloud-march throughput", "value": F * W * H / ms / 1e3, "unit": "Mpix/s", "n_gpus": world, "steps": len(times),
                          "warmup": max(1, args.warmup // 3), "ms_per_step": ms, "ms_per_frame": ms / F, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": len(times) * len(mine), "clocks": clocks,
                          "config": {"workload": f"{args.config} {W}x{H} x {F}-frame wind animation (sky.wind.w = 8k), frame-parallel: frame k on rank k % {world}",
                                     "filter": args.filter, "l2": "frames are 531 MB at 8K (> L2); no flush", "parallelism": f"frame-parallel x{world}, frames stored into rank 0's ring over NVLink"}}))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    ring.close()
    cs.close()
    if world > 1:
        dist.destroy_process_group()


def sha256_of(frame_np):
    return hashlib.sha256(np.ascontiguousarray(frame_np).tobytes()).hexdigest()


def time_frames(torch, mm, cs, stream, flush, n, warm=2, **dispatch_kw):
    """mean CUDA-event ms of n single-GPU MM_FULL frames, L2 flushed before each, events on the launching stream"""
    for _ in range(warm):
        flush.fill_(1)
        cs.dispatch(mm.MM_FULL, stream=stream.cuda_stream, **dispatch_kw)
    evs = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        cs.dispatch(mm.MM_FULL, stream=stream.cuda_stream, **dispatch_kw)
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / n


def sampler_tail(mm, ob, cs, sc):
    """The default mode (texture-unit filtering) against the BINARY32-sampler oracle on the whole frame: the tail of pixels that
    differ between the two sampler definitions.  Counters on for one frame to count flipped rays."""
    W, H = sc["W"], sc["H"]
    t0 = time.time()
    ref, rcnt = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ob.OM_FILTER_FP32).march(W, H, nthreads=host_threads())
    dt = time.time() - t0
    cs.enableCounters(True)
    cs.setFilterMode(mm.MM_FILTER_HW)
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cnt = cs.readCounters()
    cs.enableCounters(False)
    rep = ob.parity_report(ref, img, rcnt, cnt)
    a, b = ob.tonemap_rgba8(ref).astype(np.int32), ob.tonemap_rgba8(img).astype(np.int32)
    d = np.abs(a - b).max(axis=-1)
    return {"against": "oracle with the binary32 sampler (OM_FILTER_FP32), full frame", "pixels": W * H, "max_abs_diff_8bit": rep["max_abs_diff_8bit"],
            "frac_within_1": rep["frac_within_1"], "frac_within_2": float((d <= 2).mean()), "pixels_beyond_2": int((d > 2).sum()),
            "branch_flip_pixels": rep["branch_flip_pixels"], "branch_flip_frac": rep["branch_flip_pixels"] / (W * H),
            "passes_literal_gate": bool(rep["pass"]), "oracle_seconds": dt,
            "note": "MM_FILTER_HW is bit-exact against the oracle's texture-unit model (tests); this is the distance between the two sampler DEFINITIONS"}


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
    sched = {"auto": mm.MM_SCHED_AUTO, "static": mm.MM_SCHED_STATIC, "persistent": mm.MM_SCHED_PERSISTENT}[args.scheduler]

    def new_shader(scene):
        s = mm.ComputeShader(local, (scene["W"], scene["H"]), placement=scene["textures"]["placement"], curl=scene["textures"]["curl"],
                             lowRes=scene["textures"]["lowres"], hiRes=scene["textures"]["hires"])
        s.setFilterMode(fmode)
        s.setArithmetic(mm.MM_ARITH_FMA if args.arith == "fma" else mm.MM_ARITH_IEEE)
        s.setLanesPerRay(args.lanes)
        s.setScheduler(sched, args.refill)
        s.updateUniformBuffers(scene["cam"], None, scene["sky"], scene["sun"])
        return s

    cs = new_shader(sc)
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
    per_rank_mean = float(per.mean())
    if world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)      # frame time = slowest rank
    ms = float(per.mean())
    mpix = W * H / ms / 1e3
    rank_ms = [per_rank_mean]
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, per_rank_mean)
        rank_ms = box

    # ---- the frame itself: hash of rank 0's device image at every N; at N > 1 rank 0 re-renders the frame alone (second context,
    # same modes) and the sharded frame must equal it bit for bit
    frame_hash, sharded_ok = None, None
    if rank == 0:
        dev_frame = cs.readOutput()
        frame_hash = sha256_of(dev_frame)
        if world > 1:
            solo = new_shader(sc)
            solo.allocOutput()
            solo.dispatch(mm.MM_FULL)
            solo.synchronize()
            alone = solo.readOutput()
            solo.close()
            sharded_ok = bool(np.array_equal(alone.view(np.uint32), dev_frame.view(np.uint32)))
            assert sharded_ok, "the sharded frame differs from the frame rank 0 renders alone"
            del alone

    # ---- e2e: the call a user makes with HOST buffers (uniforms in, image out), pinned host memory
    e2e = None
    n_e2e = max(3, K // 2)
    if world > 1:
        # one page-locked shared-memory frame mapped by every rank: each kernel stores its pixels to rank 0's device image over
        # NVLink AND to this host frame over its own PCIe link, so the D2H transfer is fused into the march on all N GPUs
        hostframe = multigpu.SharedHostFrame(cs, rank, world, dist)

        def e2e_once():
            cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
            cs.dispatch(mm.MM_FULL, rank, world, args.row_block, stream=stream.cuda_stream)
            stream.synchronize()                        # this rank's pixels have landed in rank 0's image and in the host frame
            hostframe.barrier()                         # ... and so have every other rank's: the frame is complete in host memory
        for _ in range(2):
            e2e_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_once()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        host_hash = sha256_of(hostframe.array) if rank == 0 else None      # the host frame is read by the CPU
        if rank == 0:
            assert host_hash == frame_hash, "the host frame differs from rank 0's device image"
        hostframe.close()
        e2e = {"value": W * H / dt / 1e6, "unit": "Mpix/s", "ms_per_frame": dt * 1e3, "h2d_bytes_per_step": 328 * world,
               "d2h_bytes_per_step": W * H * 16, "host_frame_sha256": host_hash,
               "note": "uniform blocks from host on every rank; sharded march storing into rank 0's device image over NVLink and into one page-locked "
                       "shared host frame over each GPU's PCIe link (device->host transfer fused into the kernels); per frame one stream synchronize per rank + a "
                       "host-side shared-memory barrier of the ranks; wall clock"}
    else:
        host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
        hnp = host.numpy()
        for _ in range(2):
            cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        host_hash = sha256_of(hnp)
        assert host_hash == frame_hash, "the host frame differs from the device image"
        e2e = {"value": W * H / dt / 1e6, "unit": "Mpix/s", "ms_per_frame": dt * 1e3, "h2d_bytes_per_step": 328,
               "d2h_bytes_per_step": W * H * 16, "host_frame_sha256": host_hash,
               "note": "mm_render_to_host: uniform blocks from host, RGBA32F frame back to pinned host memory (stores fused into the kernel)"}
        del host

    # ---- sustained: >= sustained_seconds of back-to-back frames (no flush, no per-frame barrier), wall clock, clocks and power
    # sampled under load.  The timed region above is short enough to stay on burst clocks; a render loop is not.
    sustained = None
    if not args.no_extras and args.sustained_seconds > 0:
        batch = max(8, int(0.25 / (ms * 1e-3)))
        barrier()
        sampler2 = ClockSampler(local) if rank == 0 else None
        if sampler2:
            sampler2.start()
        t0 = time.perf_counter()
        frames = 0
        while True:
            for _ in range(batch):
                cs.dispatch(mm.MM_FULL, rank, world, args.row_block if world > 1 else 1, stream=stream.cuda_stream)
            frames += batch
            torch.cuda.synchronize()
            go = torch.tensor([1.0 if time.perf_counter() - t0 < args.sustained_seconds else 0.0], device="cuda")
            if world > 1:
                dist.all_reduce(go, op=dist.ReduceOp.MIN)       # every rank runs the same number of batches
            if float(go) == 0.0:
                break
        barrier()
        dt = time.perf_counter() - t0
        sustained = {"ms_per_frame": dt * 1e3 / frames, "value": W * H * frames / dt / 1e6, "unit": "Mpix/s", "frames": frames, "seconds": dt,
                     "clocks": sampler2.summary() if sampler2 else None,
                     "note": f"back-to-back frames in batches of {batch} (one stream synchronize per batch), wall clock, no L2 flush"}

    extras = {}
    if world > 1 and not args.no_extras:
        # ---- the comparison path SURVEY 8e asks for: every rank marches its rows into a LOCAL image, then the bands are gathered to
        # rank 0 with grouped NCCL send/recv (torch.distributed.batch_isend_irecv = ncclGroupStart/ncclSend/ncclRecv/ncclGroupEnd) and
        # scattered into the frame -- against the fused path above, whose kernels store straight into rank 0's image over NVLink.
        local_img = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        loc = new_shader(sc)
        loc.bindOutput(local_img.data_ptr())
        rows_of = [torch.from_numpy(multigpu.owned_rows(H, r, world, args.row_block)).cuda() for r in range(world)]
        frame = torch.empty((H, W, 4), dtype=torch.float32, device="cuda") if rank == 0 else None
        bufs = [torch.empty((len(rows_of[r]), W, 4), dtype=torch.float32, device="cuda") for r in range(1, world)] if rank == 0 else None

        def gather_once():
            loc.dispatch(mm.MM_FULL, rank, world, args.row_block, stream=stream.cuda_stream)
            packed = local_img.index_select(0, rows_of[rank])
            if rank == 0:
                reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, bufs[r - 1], r) for r in range(1, world)])
            else:
                reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, packed, 0)])
            for q in reqs:
                q.wait()
            if rank == 0:
                frame.index_copy_(0, rows_of[0], packed)
                for r in range(1, world):
                    frame.index_copy_(0, rows_of[r], bufs[r - 1])
        for _ in range(3):
            gather_once()
        barrier()
        gev = []
        for _ in range(max(3, K // 2)):
            barrier()
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gather_once()
            e1.record(stream)
            gev.append((e0, e1))
        barrier()
        gt = torch.tensor([a.elapsed_time(b) for a, b in gev], dtype=torch.float64, device="cuda")
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        same = None
        if rank == 0:
            same = bool(np.array_equal(frame.cpu().numpy().view(np.uint32), dev_frame.view(np.uint32)))
        loc.close()
        del local_img, frame, bufs
        extras["nccl_gather_comparison"] = {"ms_per_frame": float(gt.mean()), "fused_peer_store_ms_per_frame": ms, "gathered_frame_equals_fused_frame": same,
                                            "what": "local march + index_select of the owned rows + grouped ncclSend/ncclRecv to rank 0 + index_copy into the frame, CUDA events, max over ranks"}
        # ---- BASELINE configs[1] sharded the same way (strong scaling of the small frame)
        if args.config != "C2":
            sc2 = scenes.make_scene(mm, "C2", assets)
            c2 = new_shader(sc2)
            sh2 = multigpu.SharedFrame(c2, rank, world, dist)
            for _ in range(3):
                c2.dispatch(mm.MM_FULL, rank, world, args.row_block, stream=stream.cuda_stream)
            barrier()
            ev2 = []
            for _ in range(K):
                barrier()
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                c2.dispatch(mm.MM_FULL, rank, world, args.row_block, stream=stream.cuda_stream)
                e1.record(stream)
                ev2.append((e0, e1))
            barrier()
            t2 = torch.tensor([a.elapsed_time(b) for a, b in ev2], dtype=torch.float64, device="cuda")
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            h2 = sha256_of(c2.readOutput()) if rank == 0 else None
            barrier()
            sh2.close()
            c2.close()
            extras["extra"] = {"C2": {"workload": "C2 1920x1080 noon, full-resolution march, sharded like the headline", "filter": args.filter,
                                      "ms_per_frame": float(t2.mean()), "value": 1920 * 1080 / float(t2.mean()) / 1e3, "unit": "Mpix/s", "frame_sha256": h2}}
    if world == 1 and not args.no_extras:
        import oracle_binding as ob
        # ---- the gate-meeting sampler modes on the same workload: FP32 filtering of the march samples
        grade = {}
        for name, fm in (("hw", mm.MM_FILTER_HW), ("hybrid", mm.MM_FILTER_HYBRID), ("exact", mm.MM_FILTER_EXACT)):
            cs.setFilterMode(fm)
            t = time_frames(torch, mm, cs, stream, flush, max(3, K // 4))
            grade[name] = {"ms_per_frame": t, "Mpix_s": W * H / t / 1e3,
                           "parity": "bit-exact decisions vs the oracle's texture-unit model" if name == "hw" else "bit-exact decisions vs the oracle's binary32 sampler (the literal north-star gate)"}
        # ---- the CONTRACTED arithmetic definition beside it: its own oracle, zero branch flips against it
        # (tests/test_arith_fma_gpu.py); a different definition, reported beside the uncontracted one, never instead of it
        cs.setArithmetic(mm.MM_ARITH_FMA)
        fma = {}
        for name, fm in (("hw", mm.MM_FILTER_HW), ("hybrid", mm.MM_FILTER_HYBRID), ("exact", mm.MM_FILTER_EXACT)):
            cs.setFilterMode(fm)
            t = time_frames(torch, mm, cs, stream, flush, max(3, K // 4))
            fma[name] = {"ms_per_frame": t, "Mpix_s": W * H / t / 1e3}
        fma["parity"] = "bit-exact decisions vs oracle/cloud_march_oracle_fma.c (pinned to the reference's shader text compiled under the same lexical contraction rule)"
        grade["arith_fma"] = fma
        cs.setArithmetic(mm.MM_ARITH_FMA if args.arith == "fma" else mm.MM_ARITH_IEEE)
        cs.setFilterMode(fmode)
        try:
            grade["hw_vs_binary32_oracle"] = sampler_tail(mm, ob, cs, sc)
        except Exception as e:
            grade["hw_vs_binary32_oracle"] = {"unavailable": f"{type(e).__name__}: {e}"}
        cs.setFilterMode(fmode)
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
        extras["parity_grade"] = grade

        # ---- scheduler variants on the same workload (scheduling only; frames are bit-identical, tests/test_scheduler_gpu.py)
        schedv = {}
        for label, sv, rf in (("static_grid_K1", mm.MM_SCHED_STATIC, 0), ("persistent_K1p_tile", mm.MM_SCHED_PERSISTENT, 32),
                              ("persistent_K1p_refill16", mm.MM_SCHED_PERSISTENT, 16), ("persistent_K1p_refill8", mm.MM_SCHED_PERSISTENT, 8),
                              ("packed_two_rays_per_thread_K1x2", mm.MM_SCHED_PACKED, 0)):
            cs.setScheduler(sv, rf)
            schedv[label] = time_frames(torch, mm, cs, stream, flush, max(3, K // 4))
        cs.setScheduler(sched, args.refill)
        extras["scheduler_ms_per_frame"] = schedv

        # ---- BASELINE configs[1] beside the headline: C2 1920x1080 noon
        if args.config != "C2":
            sc2 = scenes.make_scene(mm, "C2", assets)
            c2 = new_shader(sc2)
            c2.allocOutput()
            t2 = time_frames(torch, mm, c2, stream, flush, K)
            blk = {"workload": "C2 1920x1080 noon, full-resolution march", "filter": args.filter, "ms_per_frame": t2, "value": 1920 * 1080 / t2 / 1e3, "unit": "Mpix/s",
                   "frame_sha256": sha256_of(c2.readOutput())}
            if not args.no_cpu_baseline:
                wq2 = workload_Q(ob, sc2, 7, args.filter)
                peaks2, _ = measured_peaks()
                tp = N_SM * TEXPEAK_QUADS_PER_CLK_PER_SM * peaks2["sm_max_mhz"] * 1e6
                blk["roofline_frac"] = wq2["Q"] * 1920 * 1080 / (t2 * 1e-3) / tp
                blk["Q_quads_per_pixel"] = wq2["Q"]
                try:
                    blk["hw_vs_binary32_oracle"] = sampler_tail(mm, ob, c2, sc2)
                except Exception as e:
                    blk["hw_vs_binary32_oracle"] = {"unavailable": f"{type(e).__name__}: {e}"}
            c2.close()
            extras["extra"] = {"C2": blk}

        # ---- the reference's own per-frame cadence (VulkanApplication.cpp:1053-1071): reproject the previous image, then ONE
        # compute-clouds dispatch = 1/16 of the pixels (phase k % 16), ping-pong.  Reported beside the headline, which marches
        # EVERY pixel every frame (16 reference dispatches' worth of rays).
        ping = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        pong = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        cad = new_shader(sc)
        cad.bindOutput(ping.data_ptr())
        cad.dispatch(mm.MM_FULL, stream=stream.cuda_stream)
        cur, prev = pong, ping
        k_rep, k_march = [], []
        for i in range(Wm + K):
            sun_i = sc["sun"].copy()
            sun_i[11] = float(i % 16)                                   # sun.color.a carries the pixel phase (CC:292)
            cad.updateUniformBuffers(sc["cam"], sc["cam"], sc["sky"], sun_i)
            cad.bindOutput(cur.data_ptr())
            cad.bindPrevious(prev.data_ptr())
            flush.fill_(1)
            cad.dispatchReproject(stream=stream.cuda_stream)
            r_ms = cad.lastKernelMs()
            cad.dispatch(mm.MM_PHASE16, stream=stream.cuda_stream)
            m_ms = cad.lastKernelMs()
            if i >= Wm:
                k_rep.append(r_ms)
                k_march.append(m_ms)
            cur, prev = prev, cur
        # the same pass with a camera that MOVED since the previous frame (the case reprojection exists for): every pixel's ten taps
        # select different source pixels, so the single-load path of K5 does not apply
        cfg3 = scenes.CONFIGS[args.config]
        moved = mm.host_camera((cfg3["pos"][0] - 3.0, cfg3["pos"][1], cfg3["pos"][2] - 2.0), cfg3["yaw"] - 0.004, cfg3["pitch"] + 0.002, 45.0, 1920.0 / 1080.0)
        cad.updateUniformBuffers(sc["cam"], moved, sc["sky"], sc["sun"])
        k_mov = []
        for i in range(max(3, K // 2) + 1):
            flush.fill_(1)
            cad.dispatchReproject(stream=stream.cuda_stream)
            if i:
                k_mov.append(cad.lastKernelMs())
        torch.cuda.synchronize()
        cad.close()
        del ping, pong
        kr, km = sum(k_rep) / K, sum(k_march) / K
        extras["reference_cadence"] = {"ms_per_frame": kr + km, "frames_per_s": 1e3 / (kr + km), "reproject_kernel_ms": kr, "phase16_march_kernel_ms": km,
                                       "reproject_kernel_ms_moving_camera": sum(k_mov) / len(k_mov), "launches_per_frame": 2,
                                       "note": "reference cadence: reprojection of the previous image + one MM_PHASE16 dispatch (1/16 of the pixels marched), ping-pong images, "
                                               "L2 flushed between frames; ms_per_frame = sum of the two kernels' CUDA-event times"}

        # ---- texture-pipe ceiling measured live: L1-resident filtered fetches
        try:
            ms2, q2 = cs.measureTexPeak(mm.MM_TEX_CURL, 4096)
            ms3, q3 = cs.measureTexPeak(mm.MM_TEX_HIRES, 4096)
            extras["tex_peak"] = {"bilinear_2d_Gquad_s": q2 / 1e9, "trilinear_3d_Gquad_s": q3 / 1e9, "ms": [ms2, ms3],
                                  "how": "mm_measure_tex_peak: 148 x 8 blocks x 256 threads x 4096 filtered fetches from an L1-resident texture, CUDA events"}
        except Exception as e:                      # the ceiling is supplementary: never let it break the bench line
            extras["tex_peak"] = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peaks, which = measured_peaks()
        import oracle_binding as ob
        out = {
            "metric": "cloud-march throughput", "value": mpix, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms, "ms_per_frame": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} {W}x{H} full-resolution cloud march (every pixel), shipped CloudPlacement/CurlNoiseFBM/128^3/32^3 textures",
                       "filter": args.filter, "arithmetic": args.arith, "lanes_per_ray": args.lanes or "per dispatch", "scheduler": args.scheduler + (f", refill {args.refill}" if args.refill else ""),
                       "l2": "flushed (256 MB write between frames)", "parallelism": f"row-cyclic x{world}, row block {args.row_block}" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": K, "e2e": e2e, "frame_sha256": frame_hash, "per_rank_kernel_ms": rank_ms,
        }
        if world > 1:
            out["sharded_equals_single_gpu"] = sharded_ok
        if sustained:
            out["sustained"] = sustained
        for k in ("parity_grade", "scheduler_ms_per_frame", "extra", "reference_cadence", "nccl_gather_comparison"):
            if k in extras:
                out[k] = extras[k]
        if not args.no_cpu_baseline:
            rows_step = max(1, int(W * H / 300e3))
            wq = workload_Q(ob, sc, rows_step, args.filter, args.arith)
            texpeak = N_SM * TEXPEAK_QUADS_PER_CLK_PER_SM * peaks["sm_max_mhz"] * 1e6
            achieved = wq["Q"] * W * H / (ms * 1e-3) / world           # per GPU: every rank marches 1/world of the frame's quads in the frame time
            out["roofline"] = {"bound": "tex", "achieved": achieved / 1e9, "peak": texpeak / 1e9, "unit": "Gquad/s", "frac": achieved / texpeak,
                               "traffic": None, "per_gpu": True, "Q_quads_per_pixel": wq["Q"], "loop_trips_per_pixel": wq["trips"], "lit_steps_per_pixel": wq["lit"],
                               "peak_source": f"148 SM x 4 bilinear quads/clk x sm_max_mhz ({which} clock), ONE GPU; achieved = algorithmic quads of the frame / frame time / n_gpus",
                               "hbm_floor_ms": W * H * 16 / (peaks["hbm_gbs"] * 1e9) * 1e3 / world}
            if "tex_peak" in extras:
                out["roofline"]["measured_tex_peak"] = extras["tex_peak"]
                best = max(extras["tex_peak"].get("bilinear_2d_Gquad_s", 0.0), extras["tex_peak"].get("trilinear_3d_Gquad_s", 0.0))
                if best > 0:
                    out["roofline"]["frac_of_measured_peak"] = achieved / 1e9 / best
            # second roofline: the march is FP32 / instruction-issue bound.  Warp-instructions and DRAM bytes per launch of
            # this exact command come from the committed ncu capture (profiles/); peak = 148 SM x 4 schedulers x f_SM.
            prof = os.path.join(ROOT, "profiles", f"r02_{args.config}_{args.filter}" + ("_fma" if args.arith == "fma" else "") + ".summary.csv")
            if os.path.exists(prof) and world == 1:
                rows = {l.split(",")[0]: l.strip().split(",") for l in open(prof) if l.count(",") >= 2}
                unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                try:        # DRAM bytes of one launch of this command, from the committed ncu --set full capture
                    out["roofline"]["traffic"] = sum(float(rows[k][2]) * unit[rows[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    out["roofline"]["traffic_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, " + os.path.relpath(prof, ROOT) + ")"
                except (KeyError, ValueError):
                    pass
                winst = [float(l.split(",")[2]) for l in open(prof) if l.startswith("smsp__inst_executed.sum,")]
                if winst:
                    peak_issue = N_SM * 4 * peaks["sm_max_mhz"] * 1e6
                    out["roofline_issue"] = {"bound": "instruction issue (FP32)", "achieved": winst[0] / (ms * 1e-3) / 1e9, "peak": peak_issue / 1e9,
                                             "unit": "Gwarp-inst/s", "frac": winst[0] / (ms * 1e-3) / peak_issue,
                                             "warp_instructions_per_frame": winst[0], "source": os.path.relpath(prof, ROOT)}
            v = wq["pixels"] / wq["seconds"] / 1e6
            out["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": host_threads(), "kind": "port",
                                   "sample": f"every {rows_step}th row of the same frame ({wq['pixels']} px), oracle port with counters, OpenMP, {host_threads()} threads requested explicitly",
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
