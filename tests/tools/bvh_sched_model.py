"""Offline model of k_bvh on config 4 (no GPU needed): where do the 84 ms go, and which policy would cut them?

k_bvh gives every lane one query at a time from a work counter.  A query is a chain of dependent rounds: set-up,
bounding-volume rounds (both children of a node) and leaf tests; each scheduling round the 32 lanes of a warp vote
and the warp runs ONE of the three phases (hfb_bvh.cuh, bvh_vote), the other lanes sit the round out.  The host
build of the device code (tests/emu) records every query's exact phase sequence; emu_bvh_sched_sim replays the
warps of the kernel (148 SMs x 4 blocks x 2 warps, 32 lanes each) as a discrete-event model in which a round of
phase p costs cost[p] of its warp's time.  The three costs are calibrated on the two B200 measurements of round 1
(100 k queries: 84 ms, 400 k: 194 ms); the model then answers what other hand-out orders and voting rules give.

    python tests/tools/bvh_sched_model.py [--n 100000]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tests.common import EmuScene, P, _ptr  # noqa: E402
from hppfcl_b200 import workloads as W  # noqa: E402
from hppfcl_b200.engine import build_bvh_obbrss  # noqa: E402

WARPS = 148 * 4 * 2


def traces(n):
    w = W.config4_mesh_vs_capsules(n)
    emu = EmuScene()
    nodes = build_bvh_obbrss(w["verts"], w["tris"])
    bid = emu.register_bvh_obbrss(nodes, w["verts"], w["tris"])
    hm = emu.register_shapes(P.make_shapes([P.BV_OBBRSS], [[0, 0, 0]], data=[bid]))
    hc = emu.register_shapes(w["capsules"])
    h1 = np.full(n, hm[0], dtype=np.uint32)
    h2 = np.ascontiguousarray(hc[w["hc"]])
    L = emu.L
    L.emu_bvh_trace_distance.restype = C.c_long
    L.emu_bvh_trace_distance.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 6 + [C.c_size_t, C.c_void_p]
    cap = 400 * n
    states = np.zeros(cap, dtype=np.uint8)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    req = P.DistanceRequestPOD()
    tf1, tf2 = np.ascontiguousarray(w["tf_mesh"]), np.ascontiguousarray(w["tf_caps"])
    tot = L.emu_bvh_trace_distance(emu.h, n, _ptr(h1), _ptr(tf1), _ptr(h2), _ptr(tf2), C.byref(req), _ptr(states), cap,
                                   _ptr(offsets))
    assert tot > 0
    key = np.linalg.norm(w["tf_caps"]["T"] - w["tf_mesh"]["T"], axis=1)  # a key a device pre-pass could compute
    return (L, states[:tot].copy(), offsets, key)


def sim(tr, cost, order=None, quorum=6, policy=0, warps=WARPS, slots=1):
    L, states, offsets, _ = tr
    L.emu_bvh_sched_sim.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p]
    stats = np.zeros(8)
    c = np.ascontiguousarray(cost, dtype=np.float64)
    o = None if order is None else np.ascontiguousarray(order, dtype=np.uint32)
    rc = L.emu_bvh_sched_sim(len(offsets) - 1, _ptr(states), _ptr(offsets), _ptr(o) if o is not None else None, warps,
                             _ptr(c), quorum, policy, slots, _ptr(stats))
    assert rc == 0
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    a = ap.parse_args()
    if a.n != 100000:
        print("note: the round costs are fitted to the measurements at 100 000 and 400 000 queries; with --n %d the"
              " milliseconds below are not predictions, only the round and lane counts are meaningful" % a.n)
    t1, t4 = traces(a.n), traces(4 * a.n)
    ln = np.diff(t1[2].astype(np.int64))
    print("rounds per query: mean %.0f, p99 %.0f, max %d; steps: %s" % (
        ln.mean(), np.percentile(ln, 99), ln.max(), dict(zip(("init", "bv", "leaf"), np.bincount(t1[1], minlength=5)[2:5]))))
    unit = np.array([1.0, 1.0, 1.0])
    s1, s4 = sim(t1, unit), sim(t4, unit)
    # 1. independent warps (a round costs its warp a fixed time): the longest query alone would set the time
    print("independent-warp model: makespan/longest-query = %.2f (100k), %.2f (400k): it predicts 400k/100k = %.2f,"
          " measured 194/84 = 2.31 -> the warps are NOT independent" % (
              s1[0] / ln.max(), s4[0] / np.diff(t4[2].astype(np.int64)).max(), s4[0] / s1[0]))
    # 2. throughput law: time = (bv_rounds * c_bv + leaf_rounds * c_leaf) / SMs + tail
    A = np.array([[s1[2], s1[3]], [s4[2], s4[3]]]) / 148.0
    tail = 14.0
    cb, cl = np.linalg.solve(A, np.array([84.0 - tail, 194.0 - tail]))
    print("throughput law  T = (bv_rounds * %.2f us + leaf_rounds * %.2f us) / 148 SMs + %.0f ms tail  fits both"
          " measurements (100k: %d + %d rounds, %.1f / %.1f lanes per round; 400k: %d + %d, %.1f / %.1f)" % (
              cb * 1e3, cl * 1e3, tail, s1[2], s1[3], s1[5] / s1[2], s1[6] / s1[3], s4[2], s4[3], s4[5] / s4[2], s4[6] / s4[3]))

    def predict(st):
        return (st[2] * cb + st[3] * cl) / 148.0 + tail

    for name, tr in (("100k", t1), ("400k", t4)):
        lnq = np.diff(tr[2].astype(np.int64))
        print("%s: as built %.0f ms" % (name, predict(sim(tr, unit))))
        for label, kw in (("longest first (perfect knowledge)", dict(order=np.argsort(-lnq, kind="stable"))),
                          ("nearest to the mesh centre first", dict(order=np.argsort(tr[3], kind="stable"))),
                          ("same in 8 buckets", dict(order=np.argsort(np.minimum((tr[3] / 0.5).astype(int), 7), kind="stable"))),
                          ("bounding volumes win ties", dict(policy=1)),
                          ("set-up quorum 1", dict(quorum=1)),
                          ("set-up quorum 16", dict(quorum=16)),
                          ("4 warps per SM", dict(warps=WARPS // 2)),
                          ("2 queries per lane", dict(slots=2)),
                          ("3 queries per lane", dict(slots=3)),
                          ("4 queries per lane", dict(slots=4)),
                          ("2 warps per SM", dict(warps=WARPS // 4)),
                          ("1 warp per SM", dict(warps=WARPS // 8)),
                          ("2 queries per lane, 4 warps per SM", dict(slots=2, warps=WARPS // 2)),
                          ("centre first + quorum 1 + 4 warps per SM", dict(order=np.argsort(tr[3], kind="stable"), quorum=1,
                                                                            warps=WARPS // 2)),
                          ("centre first + quorum 1 + 2 warps per SM", dict(order=np.argsort(tr[3], kind="stable"), quorum=1,
                                                                            warps=WARPS // 4))):
            st = sim(tr, unit, **kw)
            print("   %-44s %5.0f ms   (%.1f / %.1f lanes per bv / leaf round)" % (
                label, predict(st), st[5] / max(st[2], 1), st[6] / max(st[3], 1)))


if __name__ == "__main__":
    main()
