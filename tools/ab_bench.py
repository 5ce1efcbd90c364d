#!/usr/bin/env python
"""A/B timing of scheduling variants of the march on one GPU (CUDA events, L2 flushed between frames).

    python tools/ab_bench.py [--config C3] [--filter hw] [--frames 8] [--variants static,tile,refill16,...] [--shard R/N]

Variants: static (K1), tile (K1p, a tile at a time), group (K1p, block-shared 2x2 tile groups), refill16 / refill8 (K1p lane refill), lanes2/4/8 (K1s).
--shard R/N times rank R's share of an N-way row-cyclic frame (row block 8) alone -- a rank's kernel does not depend on the others.
Under ncu the per-launch metrics of each variant can be told apart by the kernel name / launch order printed here.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--filter", default="hw")
    ap.add_argument("--arith", default="ieee")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--variants", default="static,tile,group")
    ap.add_argument("--shard", default="0/1")
    ap.add_argument("--snake", action="store_true", help="boustrophedon row-cyclic partition (MM_ROWS_SNAKE)")
    ap.add_argument("--row-block", type=int, default=8)
    ap.add_argument("--size", default="", help="WxH override of the config's extent")
    ap.add_argument("--phase16", action="store_true", help="time one MM_PHASE16 dispatch (1/16 of the pixels, phase 5) instead of MM_FULL")
    ap.add_argument("--all-ranks", action="store_true", help="time every rank's share in turn and print max / mean")
    a = ap.parse_args()
    import torch
    import _pkg
    import scenes
    mm = _pkg.load_package()
    wh = {"W": int(a.size.split("x")[0]), "H": int(a.size.split("x")[1])} if a.size else {}
    sc = scenes.make_scene(mm, a.config, scenes.load_assets(), **wh)
    W, H = sc["W"], sc["H"]
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.setFilterMode({"exact": mm.MM_FILTER_EXACT, "hw": mm.MM_FILTER_HW, "hybrid": mm.MM_FILTER_HYBRID}[a.filter])
    if hasattr(cs, "setArithmetic"):
        cs.setArithmetic({"ieee": 0, "fma": 1}[a.arith])
    sun = sc["sun"].copy()
    if a.phase16:
        sun[11] = 5.0                                                    # sun.color.a carries the pixel phase (CC:292)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sun)
    mode = mm.MM_PHASE16 if a.phase16 else mm.MM_FULL
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    r0, n = (int(x) for x in a.shard.split("/"))
    table = {"static": (1, mm.MM_SCHED_STATIC, 0), "tile": (1, mm.MM_SCHED_PERSISTENT, 32), "group": (1, mm.MM_SCHED_PERSISTENT, 128),
             "packed": (1, mm.MM_SCHED_PACKED, 0), "refill16": (1, mm.MM_SCHED_PERSISTENT, 16), "refill8": (1, mm.MM_SCHED_PERSISTENT, 8),
             "lanes2": (2, mm.MM_SCHED_AUTO, 0), "lanes4": (4, mm.MM_SCHED_AUTO, 0), "lanes8": (8, mm.MM_SCHED_AUTO, 0), "auto": (0, mm.MM_SCHED_AUTO, 0)}
    for name in a.variants.split(","):
        lanes, sched, refill = table[name]
        cs.setLanesPerRay(lanes)
        try:
            cs.setScheduler(sched, refill)
        except mm.MarshmallowError as e:
            print(f"{name}: unavailable ({e})")
            continue
        ranks = range(n) if a.all_ranks else [r0]
        per_rank = []
        for r in ranks:
            ts = []
            for i in range(a.frames + 2):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                cs.dispatch(mode | (mm.MM_ROWS_SNAKE if a.snake else 0), r, n, a.row_block if n > 1 else 1, stream=stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            per_rank.append(sum(ts) / len(ts))
        if a.all_ranks:
            print(f"{a.config} {a.filter} {a.arith} {name:9s} x{n}: max {max(per_rank):.4f} ms, mean {sum(per_rank) / len(per_rank):.4f} ms, per rank {[round(x, 4) for x in per_rank]}")
        else:
            print(f"{a.config} {a.filter} {a.arith} {name:9s} shard {r0}/{n}: {per_rank[0]:.4f} ms")
    cs.close()


if __name__ == "__main__":
    main()
