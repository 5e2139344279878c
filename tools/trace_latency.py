#!/usr/bin/env python
"""Stage timeline of the one-evaluation kernel on the bench workload (vb2_llk_trace).  Run on a GPU box."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import verifybamid_b200 as vb

s = bench.make_workload("100k30x")
with vb.LLKEngine(s.problem) as eng:
    for _ in range(20):
        eng.compute_mix_llks([0.01, 0.01], [0.01, 0.01], 0.03)
    rows = []
    for rep in range(5):
        llk, st = eng.trace([0.01 + 1e-7 * rep, 0.01], [0.01, 0.01], 0.03)
        st = st.astype(np.int64)
        order = [7, 11, 1, 2, 3, 15, 4, 12, 14, 5, 6]
        cyc = st[:, order] - st[:, 0:1]
        g0 = st[:, 8] - st[:, 8].min()
        g1 = st[:, 9] - st[:, 8].min()
        print("rep %d: global start skew max %d ns, last exit %d ns after first entry" % (rep, g0.max(), g1.max()))
        names = ["first TMA issued", "tables requested", "tables (barrier)", "blob landed", "warp 0 loop done",
                 "last warp loop done", "barrier passed (*)", "combined", "bin sum stored", "partial ready", "published"]
        if rep < 4:
            continue
        for k, nme in enumerate(names):
            c = cyc[:, k]
            print("   %-18s cycles since CTA entry: min %6d  median %6d  max %6d" % (nme, c.min(), np.median(c), c.max()))
    print("info", eng.info())
    os.environ["VB2_LLK_TRACE_SESSION"] = "1"
    llk, st = eng.trace([0.01, 0.01], [0.01, 0.01], 0.03)
    st = st.astype(np.int64)
    print("session kernel, cycles since the CTA started polling for its LAST evaluation (idle wait included in the first two):")
    for k, nme in [(1, "doorbell seen"), (2, "params published"), (3, "warp 0 slices done"), (15, "last warp slices done"), (6, "published")]:
        c = st[:, k] - st[:, 0]
        print("   %-22s min %7d  median %7d  max %7d" % (nme, c.min(), np.median(c), c.max()))
    c = st[:, 6] - st[:, 2]
    print("   params published -> partial published: min %d median %d max %d cycles" % (c.min(), np.median(c), c.max()))
    c = st[:, 15] - st[:, 2]
    print("   params published -> last warp done:    min %d median %d max %d cycles" % (c.min(), np.median(c), c.max()))
    order = np.argsort(-c)
    print("   slowest CTAs:", [(int(i), int(c[i])) for i in order[:12]])
    print("   mean over CTAs 0..107: %d   108..147: %d" % (c[:108].mean(), c[108:].mean()))
    os.environ["VB2_LLK_TRACE_SESSION"] = "search"
    for rep in range(2):
        llk, st = eng.trace([0.01, 0.01], [0.01, 0.01], 0.03)
    st = st.astype(np.int64)
    print("search on the device (vb2_llk_minimize), LAST evaluation, cycles since the top of the CTA's loop:")
    for k, nme in [(1, "point + coefficients ready"), (2, "CTA barrier passed"), (3, "warp 0 slices done"), (15, "last warp slices done"),
                   (6, "partial pushed to all"), (4, "all partials gathered"), (5, "simplex stepped")]:
        c = st[:, k] - st[:, 0]
        print("   %-28s min %7d  median %7d  max %7d" % (nme, c.min(), np.median(c), c.max()))
