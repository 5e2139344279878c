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


def iter_blobs(pk: dict):
    """Yield (blob_index, bin, blob_bytes) in image order: blob q is the (q % n_bins)-th blob of round q // n_bins and
    belongs to bin first_bin + q % n_bins (the bins that own a blob in a round are a contiguous range)."""
    nb = pk["n_bins"]
    q = 0
    for r, R in enumerate(pk["rounds"]):
        for k in range(R["count"]):
            off = R["base"] + k * R["stride"]
            yield r * nb + k, R["first_bin"] + k, pk["blob"][off:off + R["stride"]]
            q += 1
    assert q == pk["n_slices"]


def emulate_packed_llk(pk: dict, pc1, pc2, alpha: float, min_af=5e-5, max_af=0.99995, kernel_form: bool = False) -> float:
    """numpy restatement of llk_kernel over the packed image (CPU check of the flatten + the maths).

    kernel_form=False: one factor F_p(e) = c0_p + c1_p e per read, one log per marker.
    kernel_form=True : the arithmetic the kernels actually run -- full rows two reads at a time through their symmetric
    functions, F_p(ea) F_p(eb) = C0_p + C1_p (ea + eb) + C2_p ea eb, the marginals of a bin multiplied up per lane
    with the exponent split off after every factor, one log per lane and bin (llk_engine.cu: eat4, combine)."""
    phred = np.power(10.0, np.arange(94) / -10.0)
    pairs, c0, c1 = job_coefficients(alpha)
    pdt = np.float64 if pk["panel_elem"] == 8 else np.float32
    k, total = pk["n_pc"], 0.0

    def gf(af):
        af = np.clip(af, min_af, max_af)
        return np.stack([(1 - af) * (1 - af), 2 * af * (1 - af), af * af])
    C0, C1, C2 = c0 * c0, c0 * c1, c1 * c1
    bin_prod, bin_exp, bin_extra = {}, {}, {}      # kernel_form: per bin, per lane running product / exponent / rare logs
    for _, bin_id, blob in iter_blobs(pk):
        wr, wa, nv_tails, full = blob[:16].view(np.uint32)
        n_valid, tail_ref, tail_alt = int(nv_tails) & 0xFF, (int(nv_tails) >> 8) & 0xF, (int(nv_tails) >> 12) & 0xF
        if pk["known_af"]:
            af1 = af2 = blob[pk["off_kaf"]:pk["off_kaf"] + 256].view(np.float64)
        else:
            ud = blob[pk["off_ud"]:pk["off_ud"] + k * 32 * pk["panel_elem"]].view(pdt).reshape(k, 32).astype(np.float64)
            mu = blob[pk["off_mu"]:pk["off_mu"] + 32 * pk["panel_elem"]].view(pdt).astype(np.float64)
            af1 = (np.asarray(pc1, dtype=np.float64) @ ud + mu) / 2.0
            af2 = (np.asarray(pc2, dtype=np.float64) @ ud + mu) / 2.0
        g1v, g2v = gf(af1), gf(af2)
        diag = blob[pk["off_diag"]:pk["off_diag"] + 768].view(np.float64).reshape(3, 32)
        byts = blob[pk["off_words"]:pk["off_words"] + (wr + wa) * 128].reshape(wr + wa, 32, 4)
        # header promise: the leading full_ref / full_alt rows carry no filler byte in any valid lane
        fr, fa = int(full) & 0xFFFF, int(full) >> 16
        assert fr <= wr and fa <= wa
        assert not (byts[:fr, :n_valid] == 0xFF).any() and not (byts[wr:wr + fa, :n_valid] == 0xFF).any()
        # ... and a uniform tail holds exactly tail_x reads (then fillers) in the row after the full rows, in every valid lane
        for tail, row in ((tail_ref, fr), (tail_alt, wr + fa)):
            if tail:
                assert 1 <= tail <= 3 and row == (wr - 1 if row == fr else wr + wa - 1)
                assert not (byts[row, :n_valid, :tail] == 0xFF).any() and (byts[row, :n_valid, tail:] == 0xFF).all()
        acc = np.ones((6, 32))
        for sect, lo, hi, n_full in (("ref", 0, wr, fr), ("alt", wr, wr + wa, fa)):
            q = byts[lo:hi].transpose(1, 0, 2).reshape(32, -1)      # [lane, reads]
            pad = q == 0xFF
            e = phred[np.where(pad, 0, q)]
            for p in range(6):
                tgt = p if sect == "ref" else 5 - p
                if kernel_form:                                      # full rows: two reads at a time
                    ef = e[:, :4 * n_full].reshape(32, -1, 2)
                    g = C0[p] + C1[p] * (ef[:, :, 0] + ef[:, :, 1]) + C2[p] * (ef[:, :, 0] * ef[:, :, 1])
                    acc[tgt] *= g.prod(axis=1)
                    rest, rpad = e[:, 4 * n_full:], pad[:, 4 * n_full:]
                    acc[tgt] *= np.where(rpad, 1.0, c1[p] * rest + c0[p]).prod(axis=1)
                else:
                    acc[tgt] *= np.where(pad, 1.0, c1[p] * e + c0[p]).prod(axis=1)
        if kernel_form:                                              # running products start at their weights
            L = sum(diag[g] * g1v[g] * g2v[g] for g in range(3)) + sum(acc[p] * g1v[a] * g2v[b] for p, (a, b) in enumerate(pairs))
        else:
            L = sum(diag[g] * g1v[g] * g2v[g] for g in range(3))
            for p, (a, b) in enumerate(pairs):
                L = L + acc[p] * g1v[a] * g2v[b]
        valid = (np.arange(32) < n_valid) & (L > 0)
        if not kernel_form:
            total += float(np.log(L[valid]).sum())
            continue
        Lv = np.where(valid, L, 1.0)
        prod = bin_prod.setdefault(bin_id, np.ones(32))
        esum = bin_exp.setdefault(bin_id, np.zeros(32, dtype=np.int64))
        extra = bin_extra.setdefault(bin_id, np.zeros(32))
        big = Lv > 1e-280
        prod *= np.where(big, Lv, 1.0)
        extra += np.where(big, 0.0, np.log(np.where(big, 1.0, Lv)))
        m, ex = np.frexp(prod)                                       # prod = m * 2^ex, m in [0.5, 1)
        prod[:] = m * 2.0
        esum += ex - 1
    if kernel_form:
        for b in sorted(bin_prod):
            total += float((bin_extra[b] + (np.log(bin_prod[b]) + bin_exp[b] * np.log(2.0))).sum())
    return total + pk["log_other_const"]
