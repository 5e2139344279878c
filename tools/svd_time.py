#!/usr/bin/env python
"""Panel construction at panel size: vb2_svd_gram on the device against the reference's own ComputeSvdGram (Eigen, all
host threads) on the same centred matrix.   python tools/svd_time.py [n_marker] [n_sample] [n_pc]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from verifybamid_b200 import svd
from oracle import svd_oracle as so

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rng = np.random.default_rng(1)
pop = rng.integers(0, 5, n)
af = np.clip(rng.uniform(0.05, 0.95, (m, 1)) + rng.normal(0, 0.12, (m, 5)), 0.01, 0.99).astype(np.float32)
g = np.zeros((m, n), np.int8)
for lo in range(0, m, 5000):
    p = af[lo:lo + 5000][:, pop]
    g[lo:lo + 5000] = (rng.random(p.shape, dtype=np.float32) < p).astype(np.int8) + (rng.random(p.shape, dtype=np.float32) < p).astype(np.int8)
print("genotype matrix %d markers x %d samples (%.0f MB int8), %d PCs" % (m, n, g.nbytes / 1e6, k))
svd.svd_gram(g[:2000], k)      # context + cuSOLVER handle warm-up
best = None
for rep in range(3):
    t0 = time.perf_counter()
    r = svd.svd_gram(g, k)
    wall = time.perf_counter() - t0
    t = r["timing"]
    print("device rep %d: centre %.1f ms, A^T*A %.1f ms (%.1f TFLOP/s fp32 over the %d lower tiles' flops), eigensolver %.1f ms, A*V %.1f ms; "
          "device total %.1f ms; call incl. H2D/D2H and allocation %.0f ms" % (rep, t["center_ms"], t["gram_ms"],
          2.0 * m * n * n / 2 / (t["gram_ms"] * 1e-3) / 1e12, (n + 127) // 128 * ((n + 127) // 128 + 1) // 2, t["eigen_ms"], t["ud_ms"], t["total_ms"], wall * 1e3))
    best = r if best is None else best
if so.reference_available():
    a, mu = so.center(g)
    assert (mu == r["mu"]).all()
    t0 = time.perf_counter()
    rud, rpc, rsv = so.reference_svd(a, k, "gram")
    ref_s = time.perf_counter() - t0
    scale = float(np.linalg.norm(rud[:, 0]))
    errs = [so.column_error(rud[:, c], r["ud"][:, c], scale) for c in range(k)]
    print("reference ComputeSvdGram (Eigen, %d threads) incl. matrix file I/O: %.1f s; worst UD column error vs it %.1e (top 3: %.1e)"
          % (os.cpu_count(), ref_s, max(errs), max(errs[:3])))
