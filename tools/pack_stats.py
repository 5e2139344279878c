#!/usr/bin/env python
"""Per-bin / per-CTA work of the bench workload's flattened image (host only): full rows, uniform tails, checked rows."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import verifybamid_b200 as vb

s = bench.make_workload()
pk = vb.pack_host(s.problem, panel_dtype=vb.VB2_PANEL_FP32)
blob = pk["blob"]
nb = pk["n_bins"]
full = np.zeros(nb); tails = np.zeros(nb); checked = np.zeros(nb); slices = np.zeros(nb)
tot = {"slices": 0, "with_checked": 0, "checked_rows": 0, "tail_rows": 0, "full_rows": 0}
for R in pk["rounds"]:
    for k in range(R["count"]):
        b = R["first_bin"] + k
        off = R["base"] + k * R["stride"]
        wr, wa, z, w = np.frombuffer(blob[off:off + 16].tobytes(), dtype=np.uint32)
        fr, fa = int(w) & 0xFFFF, int(w) >> 16
        tr, ta = (int(z) >> 8) & 0xF, (int(z) >> 12) & 0xF
        rr, ra = int(wr) - fr, int(wa) - fa
        nt = (1 if (tr and rr == 1) else 0) + (1 if (ta and ra == 1) else 0)
        nc = (0 if (tr and rr == 1) else rr) + (0 if (ta and ra == 1) else ra)
        full[b] += fr + fa; tails[b] += nt; checked[b] += nc; slices[b] += 1
        tot["slices"] += 1; tot["with_checked"] += nc > 0; tot["checked_rows"] += nc; tot["tail_rows"] += nt; tot["full_rows"] += fr + fa
print(tot)
for name, cost in (("rows only", full + tails + checked), ("pack cost model (4 full, 2 tail, 7 checked, 15 per slice)", 4 * full + 2 * tails + 7 * checked + 15 * slices)):
    cta = cost.reshape(-1, 4).sum(axis=1)
    print("%-58s per bin max/mean %.3f   per CTA max/mean %.3f   slowest CTAs %s" % (name, cost.max() / cost.mean(), cta.max() / cta.mean(),
          list(np.argsort(-cta)[:8])))
print("checked rows per CTA (top):", sorted([(int(c), i) for i, c in enumerate(checked.reshape(-1, 4).sum(axis=1))], reverse=True)[:10])
