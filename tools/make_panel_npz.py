#!/usr/bin/env python
"""Pack a reference SVD panel (<prefix>.UD/.mu/.bed/.V, plain text) into one
compressed .npz that travels with the repository (the GPU box has no /root/reference).

    python tools/make_panel_npz.py /root/reference/resource/1000g.phase3.100k.b37.vcf.gz.dat \
           verifybamid_b200/data/1000g.phase3.100k.b37.npz

Values are kept as float64 exactly as `operator>>` parses them, so a panel expanded back to text by
verifybamid_b200.panels.write_text_panel() parses to the same doubles.
"""
import sys
import numpy as np


def main(prefix: str, out: str) -> None:
    ud = np.loadtxt(prefix + ".UD", dtype=np.float64, ndmin=2)
    names, mu = [], []
    with open(prefix + ".mu") as f:
        for line in f:
            a, b = line.split()[:2]
            names.append(a); mu.append(float(b))
    chrom, pos, ref, alt = [], [], [], []
    with open(prefix + ".bed") as f:
        for line in f:
            t = line.split()
            chrom.append(t[0]); pos.append(int(t[2])); ref.append(t[3]); alt.append(t[4])
    vid, v = [], []
    with open(prefix + ".V") as f:
        for line in f:
            t = line.split()
            vid.append(t[0]); v.append([float(x) for x in t[1:]])
    np.savez_compressed(out, ud=ud, mu=np.asarray(mu), chrom=np.asarray(chrom),
                        pos=np.asarray(pos, dtype=np.int64), ref=np.asarray(ref), alt=np.asarray(alt),
                        v_ids=np.asarray(vid), v=np.asarray(v, dtype=np.float64))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
