"""One-off diagnostic for the mismatch the GPU fuzz found (seed 3, cached-guess distance case, rows 1406 and
3708: GJK hits its iteration cap in the oracle, the GPU reports another status).  Sweeps gjk_max_iterations
1..128 on those rows and stores, per cap, the records and exit rays of the device and of the oracle, plus the
device's answers for lane groups of 1, 2 and 4 threads -> gpurun_out/diag_seed3.npz"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.dirname(__file__))
import fuzz_ref as F  # noqa: E402
from fuzz_ref import P  # noqa: E402

ROWS = [1406, 3708, 3938]


def main(make_device):
    t0 = time.time()
    out = {}
    for gc in ("2", "1", "4"):
        os.environ["HFB_GC"] = gc
        B, tag, cases = F.build_cases(3, 4000, False, make_device())
        c = cases[2]
        if gc == "2":
            full = F.run_case(B.emu, c)
            out["full_gc2"] = full[ROWS]
            out["oracle_full"] = F.run_case(B.orc, c)[ROWS]
        h1, t1, h2, t2 = [np.ascontiguousarray(c[k][ROWS]) for k in (2, 3, 4, 5)]
        gg, gh = np.ascontiguousarray(c[8][0][ROWS]), np.ascontiguousarray(c[8][1][ROWS])
        recs, rays, orecs, orays = [], [], [], []
        for cap in range(1, 129):
            req = P.DistanceRequestPOD(gjk_convergence_criterion=P.DualityGap, gjk_convergence_criterion_type=P.Relative,
                                       gjk_initial_guess=P.CachedGuess, gjk_max_iterations=cap)
            req.q.cached_gjk_guess = gg.ctypes.data
            req.q.cached_support_func_guess = gh.ctypes.data
            r, g, _ = B.emu.batch_distance(h1, t1, h2, t2, req, want_guess=True)
            recs.append(r)
            rays.append(g)
            if gc == "2":
                r, g, _ = B.orc.batch_distance(h1, t1, h2, t2, req, want_guess=True, nthreads=1)
                orecs.append(r)
                orays.append(g)
        out["dev_rec_gc" + gc] = np.stack(recs)
        out["dev_ray_gc" + gc] = np.stack(rays)
        if gc == "2":
            out["orc_rec"] = np.stack(orecs)
            out["orc_ray"] = np.stack(orays)
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/diag_seed3.npz", **out)
    d, o = out["dev_rec_gc2"], out["orc_rec"]
    for k, row in enumerate(ROWS):
        ne = np.nonzero((d["status"][:, k] != o["status"][:, k]) |
                        (out["dev_ray_gc2"][:, k].view(np.uint64) != out["orc_ray"][:, k].view(np.uint64)).any(axis=1))[0]
        print("row", row, "first differing cap", (int(ne[0]) + 1) if len(ne) else None, "full-batch status dev %x oracle %x"
              % (out["full_gc2"]["status"][k], out["oracle_full"]["status"][k]))
    print("diag %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "emu":
        from tests.common import EmuScene
        main(EmuScene)
    else:
        import hppfcl_b200 as hf
        main(lambda: hf.Engine(0))
