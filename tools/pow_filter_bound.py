#!/usr/bin/env python
"""How many warp-trips of K1 would skip the deterministic pow (det_powf, 6.6 % of K1's instructions, profiles/r02final_k1_stages.txt)
if a cheap estimate of coverage = h^k decided the outcome of CC:247-250 whenever it can?  Measured with the oracle, at WARP granularity.

The oracle classifies the coverage pow of every march trip (oracle/cloud_march_oracle.c, pow_filter_class): 0 = no pow (coverage <= 0.7,
k == 1), 1 = the estimate proves coverage > erosion FBM (result = gate density), 2 = the estimate proves the result is +0, 3 = the exact
pow is needed.  Every prediction is compared bit for bit with the exact result (om_powclass_mismatches must stay 0).  K1 is
warp-synchronous -- lane i's n-th trip runs with lane j's n-th trip -- so a warp pays for the out-of-line pow in iteration n when ANY lane of
its 8x4 pixel tile calls it; with the filter, when any lane is class 3.  Rows: every `stride`-th block of 4 rows at the config's full width
(tiles keep their true angular size).

    python tools/pow_filter_bound.py [--config C3] [--stride 27] [--filter texunit|fp32]
"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob, scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--stride", type=int, default=27)
    ap.add_argument("--filter", default="texunit")
    a = ap.parse_args()
    assets = scenes.load_assets()
    sc = scenes.scene_from_config(a.config, assets)
    W, H = sc["W"], sc["H"]
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ob.OM_FILTER_TEXUNIT if a.filter == "texunit" else ob.OM_FILTER_FP32)
    lib = ob.lib()
    rows = [y for y in range(H) if (y // 4) % a.stride == 0]
    # the oracle indexes the buffer by absolute pixel; keep it small by marching a frame that holds only the owned rows' band
    buf = np.zeros((H, W, 256), np.uint8) if H * W * 256 < (6 << 30) else None
    if buf is None:
        sys.exit("frame too large for the per-pixel class buffer; lower the resolution")
    lib.om_set_powclass_buffer.argtypes = [C.c_void_p]
    lib.om_set_powclass_buffer(buf.ctypes.data)
    _, cnt = S.march(W, H, row_begin=0, row_stride=a.stride, row_block=4, nthreads=os.cpu_count())
    lib.om_set_powclass_buffer(None)
    mism = C.c_ulonglong.in_dll(lib, "om_powclass_mismatches").value
    top = buf[rows] >> 7                                      # bit 7: the trip's height is above every height gradient (h >= 0.901)
    cls = buf[rows] & 0x7f                                    # [owned rows][W][256]
    trips = cnt[rows][..., 0]
    nrows = len(rows) // 4 * 4
    cls = cls[:nrows].reshape(nrows // 4, 4, W // 8, 8, 256).transpose(0, 2, 1, 3, 4).reshape(-1, 32, 256)      # [tile][lane][trip]
    trips = trips[:nrows].reshape(nrows // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
    assert trips.max() <= 256
    warp_trips = int(trips.max(axis=1).sum())
    lane_trips = int(trips.sum())
    calls = cls > 0
    print(f"{a.config} {W}x{H}, {len(rows)} rows, sampler {a.filter}: {lane_trips:,} lane-trips in {warp_trips:,} warp-trips "
          f"({lane_trips / warp_trips:.1f} live lanes per warp-trip); prediction mismatches {mism}")
    n = calls.sum()
    print(f"  lane level: pow calls on {n / lane_trips:.3f} of trips; of those class 1 (coverage > FBM) {np.mean(cls[calls] == 1):.3f}, "
          f"class 2 (provably +0) {np.mean(cls[calls] == 2):.3f}, class 3 (exact pow needed) {np.mean(cls[calls] == 3):.3f}")
    pay_now = calls.any(axis=1)                               # [tile][trip]
    pay_filtered = (cls == 3).any(axis=1)
    lanes_now = calls.sum(axis=1)[pay_now].mean()
    print(f"  warp level: {pay_now.sum() / warp_trips:.3f} of warp-trips call the pow today ({lanes_now:.1f} lanes in it on average); "
          f"with the filter {pay_filtered.sum() / warp_trips:.3f} -> {1 - pay_filtered.sum() / max(1, pay_now.sum()):.3f} of the calls avoided")
    # Second question, same trace: once a ray is above every height gradient (h >= 0.901: cumulus, stratocumulus and stratus are all exactly 0) it stays
    # there -- the distance from the earth's centre grows along an upward ray -- and every remaining trip is a miss whose only effects are the counters of
    # CC:468-481 and t += stepSize.  A warp whose live lanes have all been SEEN in that zone on an earlier trip could run those trips without the ~150
    # instructions of position / height / gradients.  How many warp-trips is that?
    top = top[:nrows].reshape(nrows // 4, 4, W // 8, 8, 256).transpose(0, 2, 1, 3, 4).reshape(-1, 32, 256)
    n_idx = np.arange(256)[None, None, :]
    live = n_idx < trips[:, :, None]                          # lane executes trip n
    seen_before = (np.cumsum(top, axis=2) - top) > 0          # flagged on an earlier trip
    assert not (seen_before & live & (top == 0)).any(), "a ray left the top zone again"
    skippable = live.any(axis=1) & ~(live & ~seen_before).any(axis=1)
    print(f"  top zone (h >= 0.901): {(top.astype(bool) & live).sum() / lane_trips:.3f} of lane-trips lie in it; warp-trips in which EVERY live lane was already seen there "
          f"(skippable without computing a height): {skippable.sum() / warp_trips:.3f}")
    assert mism == 0, "the filter mispredicted an outcome"


if __name__ == "__main__":
    main()
