#!/usr/bin/env python
"""Time of the stage in front of the kernels on the headline workload: pileup text -> resident engine, on the device
(vb2_ingest_*) vs on the host (C++ reader + BuildResolvedMarkers + host flatten + upload).  Run on a GPU box."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import verifybamid_b200 as vb
from verifybamid_b200 import host, panels

s = bench.make_workload("100k30x")
with tempfile.TemporaryDirectory() as td:
    prefix = panels.write_text_panel(s.panel, os.path.join(td, "panel"))
    pile = s.write_pileup(os.path.join(td, "sample.pileup"))
    text = open(pile, "rb").read()
    dp = vb.DevicePanel(s.panel.ud[:, :2], s.panel.mu, s.panel.chrom, s.panel.pos, s.panel.alt_char())
    dev, hst = [], []
    for rep in range(6):
        t0 = time.perf_counter()
        eng = dp.ingest(text)
        dev.append(time.perf_counter() - t0)
        eng.close()
        t0 = time.perf_counter()
        prob, _ = host.load_problem(prefix, pile, 2, disable_sanity=False)   # (also re-reads the panel text: subtracted below)
        t1 = time.perf_counter()
        e2 = vb.LLKEngine(prob)
        t2 = time.perf_counter()
        e2.close()
        hst.append((t1 - t0, t2 - t1))
    print("pileup text: %.1f MB, %d lines" % (len(text) / 1e6, text.count(b"\n")))
    print("device: text -> resident engine  %s ms (best %.2f)" % (["%.2f" % (x * 1e3) for x in dev], min(dev) * 1e3))
    print("host  : panel + pileup read %s ms, flatten + upload %s ms" % (["%.1f" % (a * 1e3) for a, _ in hst], ["%.1f" % (b * 1e3) for _, b in hst]))
    dp.close()
