"""Offline look at the GJK kernel of config 2 (no GPU needed): how full are its warps, and what would passes over
iteration buckets give?

k_pairs<1, CAP_PRIM, GJK route> gives a thread one pair of the class-sorted list per trip; a warp's trip lasts as long
as its slowest pair: max over 32 lanes of the GJK iteration count (exact, from the oracle).  The script reports the
lane utilisation of that loop and replays two alternatives on the same counts:
  * passes: every pair runs at most b1 iterations, the survivors are compacted (order kept) and run b2 more, ...
    (one state save/restore per surviving pair and pass);
  * an oracle-sorted order (pairs sorted by iteration count inside a class): the bound of any re-binning.

    python tests/tools/gjk_sched_model.py [--n 1000000]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hppfcl_b200 import _pod as P, workloads as W  # noqa: E402
from oracle import oracle_lib  # noqa: E402


def warp_rounds(iters):
    """sum over trips of the slowest lane; `iters` in hand-out order"""
    n = len(iters)
    pad = (-n) % 32
    x = np.concatenate([iters, np.zeros(pad, dtype=iters.dtype)]).reshape(-1, 32)
    return int(x.max(axis=1).sum()), int(x.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    a = ap.parse_args()
    w = W.config2_mixed_primitives(a.n)
    orc = oracle_lib.OracleScene(P)
    h = orc.register_shapes(w["shapes"])
    r = orc.batch_distance(h[w["h1"]], w["tf1"], h[w["h2"]], w["tf2"], nthreads=0)
    gjk = P.status_path(r["status"]) == P.PATH_GJK
    it = (r["iterations"] & 0xffff).astype(np.int64)
    t1, t2 = w["shapes"]["type"][w["h1"]], w["shapes"]["type"][w["h2"]]
    cls = t1.astype(np.int64) * 64 + t2  # the kernel's class sort: by ordered type pair
    idx = np.nonzero(gjk)[0]
    order = idx[np.argsort(cls[idx], kind="stable")]
    x = it[order]
    print("GJK-routed pairs: %d of %d; iterations mean %.1f, p50 %d, p99 %d, max %d" % (
        len(x), a.n, x.mean(), np.percentile(x, 50), np.percentile(x, 99), x.max()))
    wr, work = warp_rounds(x)
    print("as built (class-sorted, 32 consecutive pairs per warp trip): %d warp iterations for %d pair iterations: "
          "%.1f of 32 lanes busy" % (wr, work, work / wr))
    # bound: sorted by iteration count inside each class
    srt = np.concatenate([np.sort(x[cls[order] == c]) for c in np.unique(cls[order])])
    wr_s, _ = warp_rounds(srt)
    print("perfect re-binning inside a class (bound): %d warp iterations (%.0f %% of as built)" % (wr_s, 100.0 * wr_s / wr))
    for buckets in ((3, 3, 4), (4, 4), (2, 2, 2, 4), (6,)):
        rem = x.copy()
        total = 0
        moved = 0
        for b in buckets:
            run = np.minimum(rem, b)
            r_, _ = warp_rounds(run)
            total += r_
            rem = rem - run
            rem = rem[rem > 0]  # compaction keeps the order
            moved += len(rem)
        r_, _ = warp_rounds(rem)
        total += r_
        print("passes of %s + rest: %d warp iterations (%.0f %% of as built), %d state save/restores (%.2f per pair)" % (
            "+".join(map(str, buckets)), total, 100.0 * total / wr, moved, moved / len(x)))


if __name__ == "__main__":
    main()
