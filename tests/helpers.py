"""Shared test helpers.  The oracle (oracle/) is imported here and ONLY under tests/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import vb2_oracle as vo  # noqa: E402
from verifybamid_b200 import PileupProblem  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "hapmap")
HAPMAP = os.path.join(GOLD, "hapmap_3.3.b37.dat")
RESULT_PILEUP = os.path.join(GOLD, "expected", "result.Pileup")
LONGREAD_PILEUP = os.path.join(GOLD, "test.LongRead.pileup")

# known-answer values of ComputeMixLLKs (reference binary, fp64),
# notation LLK(pc_contam, pc_intended; alpha)
Z, A, B = [0.0, 0.0], [0.01, 0.01], [0.05, -0.02]
KAT_POINTS = [(Z, Z, 0.5), (A, A, 0.03), (B, A, 0.10), (A, A, 0.0), (B, A, 0.999)]
KAT_RESULT = [-21.3636239491917, -21.0164342381918, -20.31310860241, -21.0705640434662, -20.603292577323]
KAT_LONGREAD = [-392.836261382088, -394.64535081385, -394.900570522841, -394.466753849932, -397.792305304655]


def to_product(p: "vo.Problem") -> PileupProblem:
    return PileupProblem(p.ud, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, p.known_af,
                         p.sanity_disabled, p.avg_depth, p.sd_depth, p.n_marker_total)


def to_oracle(p: PileupProblem, num_thread: int = 8) -> "vo.Problem":
    return vo.Problem(p.ud, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, p.known_af,
                      p.sanity_disabled, p.avg_depth, p.sd_depth, num_thread, p.n_marker_total)


def golden_problem(pileup: str, n_pc: int = 2) -> "vo.Problem":
    return vo.problem_from_files(HAPMAP, pileup, n_pc, disable_sanity=True)


def job_coefficients(alpha: float):
    """c0/c1 of the six off-diagonal genotype pairs (llk_engine.cu fill_job)."""
    E = [0.0, 1.0 / 6.0, 1.0 / 3.0]
    N = [1.0, 0.5, 0.0]
    pairs = [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]
    c0, c1 = [], []
    for g1, g2 in pairs:
        e_mix = alpha * E[g1] + (1.0 - alpha) * E[g2]
        n_mix = alpha * N[g1] + (1.0 - alpha) * N[g2]
        c0.append(n_mix)
        c1.append(e_mix - n_mix)
    return pairs, np.array(c0), np.array(c1)


def emulate_packed_llk(pk: dict, pc1, pc2, alpha: float, min_af=5e-5, max_af=0.99995) -> float:
    """numpy restatement of llk_kernel over the packed image (CPU check of the flatten + the maths)."""
    phred = np.power(10.0, np.arange(94) / -10.0)
    pairs, c0, c1 = job_coefficients(alpha)
    m_pad, total = pk["m_pad"], 0.0
    if pk["known_af"] is not None:
        af1 = af2 = pk["known_af"]
    else:
        af1 = (np.asarray(pc1) @ pk["ud"] + pk["mu"]) / 2.0
        af2 = (np.asarray(pc2) @ pk["ud"] + pk["mu"]) / 2.0

    def gf(af):
        af = np.clip(af, min_af, max_af)
        return np.stack([(1 - af) * (1 - af), 2 * af * (1 - af), af * af])
    g1v, g2v = gf(af1), gf(af2)
    for s in range(pk["n_slices"]):
        base, wrwa = int(pk["slice_desc"][s, 0]), int(pk["slice_desc"][s, 1])
        wr, wa = wrwa & 0xFFFF, wrwa >> 16
        blk = pk["words"][base:base + (wr + wa) * 32].reshape(wr + wa, 32)
        byts = blk.view(np.uint8).reshape(wr + wa, 32, 4)          # little-endian bytes of each word
        acc = np.ones((6, 32))
        for sect, lo, hi in (("ref", 0, wr), ("alt", wr, wr + wa)):
            q = byts[lo:hi].transpose(1, 0, 2).reshape(32, -1)      # [lane, reads]
            pad = q == 0xFF
            e = phred[np.where(pad, 0, q)]
            for p in range(6):
                f = np.where(pad, 1.0, c1[p] * e + c0[p])
                tgt = p if sect == "ref" else 5 - p
                acc[tgt] *= f.prod(axis=1)
        pm = np.arange(s * 32, s * 32 + 32)
        L = sum(pk["diag"][g, pm] * g1v[g, pm] * g2v[g, pm] for g in range(3))
        for p, (a, b) in enumerate(pairs):
            L = L + acc[p] * g1v[a, pm] * g2v[b, pm]
        valid = (pm < pk["n_used"]) & (L > 0)
        total += float(np.log(L[valid]).sum())
    return total + pk["log_other_const"]
