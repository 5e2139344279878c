#!/usr/bin/env python
"""bench.py -- marker-reads/sec per LLK evaluation (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (`--config 100k30x`, the default): BASELINE.json configs[1] -- synthetic pileup on the
1000g.phase3.100k.b37 panel, 100k markers x 30x, NumPC=2, alpha=0.02.  Extra, non-headline lines:
`--config k4` (configs[2]: NumPC=4), `--config hgdp200` (configs[4]: hgdp.100k x 200x, NumPC=4, the chunked deep-
coverage path), `--config batch64` (configs[3]: 64 samples, spread over the GPUs, no collective).

A STEP is one launch of the many-evaluations kernel over EVALS_PER_STEP (2,048) independent evaluations of the
contamination log-likelihood (2,048 calls of the reference's ComputeMixLLKs, ContaminationEstimator.h:194-314); the
same at every N, so `--steps 20` times about 0.1 s of device work.  With N GPUs the markers are sharded across the
ranks (32-marker slices, round-robin), every rank evaluates its shard for the step's 2,048 parameter sets and ONE
all-reduce adds the 2,048 partial sums (overlapped with the next step's kernel) -- strong scaling of a fixed sample.

Keys of the JSON line (rank 0):
  value     marker-reads/s = reads per evaluation x EVALS_PER_STEP / step time, everything resident in HBM, device-timed
            (CUDA events on the launching stream) over exactly K steps; a step rotates through enough resident copies
            of the sample to exceed L2, so every evaluation streams its sample from HBM;
  e2e       the same metric through the public C-ABI call with HOST buffers, host<->device traffic of every evaluation
            inside the timed region.  N=1: the call the simplex search makes -- EVALS_PER_STEP DEPENDENT evaluations
            per step (vb2_llk_eval one after the other inside an evaluation session; the next one starts when the
            last one's scalar is back on the host).  N>1: the batched public call (host parameter arrays in, all-
            reduced results back in pinned host memory);
  roofline  algorithmic bytes of one launch (2*R + 4*(k+2)*M' per evaluation x evaluations per launch)
            / the launch's measured duration, against the measured HBM peak;
  parity    the step's results and the e2e result checked against the CPU oracle inside this run (relative error);
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, else the C port) on this host at 1, 4 and all threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "marker-reads/sec per LLK eval"
UNIT = "marker-reads/s"
EVALS_PER_STEP = 2048          # evaluations per step (= per launch of the many-evaluations kernel), at every N
REF_EVALS_PER_STEP = 16        # reference arm: bounded sample of a step's evaluations
ALPHA, SEED = 0.02, 1
L2_BYTES = 126 * 1024 * 1024
PARITY_TOL = 1e-8

CONFIGS = {
    # name: (panel, n_pc, depth, samples, BASELINE.json configs[] index, description)
    "100k30x": ("1000g.phase3.100k.b37", 2, 30.0, 1, 1,
                "synthetic pileup, 1000g.phase3.100k.b37 panel, 100k markers x 30x, NumPC=2, alpha=0.02, sanity filter on (seed 1)"),
    "k4": ("1000g.phase3.100k.b37", 4, 30.0, 1, 2,
           "synthetic pileup, 1000g.phase3.100k.b37 panel, 100k markers x 30x, NumPC=4, alpha=0.02, sanity filter on (seed 1)"),
    "hgdp200": ("hgdp.100k.b37", 4, 200.0, 1, 4,
                "synthetic pileup, hgdp.100k.b37 panel, 100k markers x 200x, NumPC=4, alpha=0.02, sanity filter on (seed 1)"),
    "batch64": ("1000g.phase3.100k.b37", 2, 30.0, 64, 3,
                "batch of 64 synthetic samples (seeds 1..64), 1000g.phase3.100k.b37 panel, 100k markers x 30x each, NumPC=2, "
                "alpha=0.02, sanity filter on; samples spread over the GPUs, no collective"),
}


def make_workload(name: str, seed: int = SEED):
    from verifybamid_b200 import panels, synth
    panel_name, n_pc, depth, _, _, _ = CONFIGS[name]
    panel = panels.load_bundled(panel_name)
    return synth.make_sample(panel, n_pc=n_pc, depth=depth, alpha=ALPHA, seed=seed, sanity_check=True)


def config_dict(name: str, sample, n_gpus: int) -> dict:
    """The `config` object: identical in both arms (--impl ours / reference) for the same --config and --gpus."""
    panel_name, n_pc, depth, n_samples, idx, text = CONFIGS[name]
    markers, reads = sample.problem.used_counts()
    if n_samples > 1:
        par = "%d samples over %d GPU(s), one launch per step evaluates every sample of a GPU once, no collective" % (n_samples, n_gpus)
    elif n_gpus > 1:
        par = "marker shards x%d, the step's %d shard sums added across the GPUs once per step" % (n_gpus, EVALS_PER_STEP)
    else:
        par = "single GPU"
    return {"workload": text, "baseline_config_index": idx, "name": name, "reads_per_eval": reads, "markers_used": markers,
            "n_pc": n_pc, "evals_per_step": EVALS_PER_STEP if n_samples == 1 else n_samples, "parallelism": par}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on this host's cores
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU path on one workload: oracle/_ref/vb2_ref (the reference's own sources, built by
    oracle/Makefile) when present, else the C port of the oracle.  Test/bench infrastructure only."""

    def __init__(self, sample):
        from oracle import vb2_oracle as vo  # checker / baseline only -- never on the product path
        self.vo, self.sample, self.p = vo, sample, sample.problem
        self.kind = "reference" if vo.ref_available() else "port"
        self.td = None
        self.n_pc = self.p.n_pc

    def __enter__(self):
        if self.kind == "reference":
            from verifybamid_b200 import panels
            self.td = tempfile.TemporaryDirectory()
            self.prefix = panels.write_text_panel(self.sample.panel, os.path.join(self.td.name, "panel"))
            self.pile = self.sample.write_pileup(os.path.join(self.td.name, "sample.pileup"))
        return self

    def __exit__(self, *exc):
        if self.td:
            self.td.cleanup()

    def oracle_problem(self, threads: int):
        p = self.p
        return self.vo.Problem(p.ud, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, None,
                               p.sanity_disabled, p.avg_depth, p.sd_depth, threads)

    def time_evals(self, evals: int, warmup: int, threads: int, converge: bool = False):
        """(seconds for `evals` full evaluations, records of the reference run or None)."""
        if self.kind == "reference":
            recs = self.vo.run_ref(["--SVDPrefix", self.prefix, "--PileupFile", self.pile, "--NumPC", str(self.n_pc),
                                    "--NumThread", str(threads), "--BenchEvals", str(evals), "--BenchWarmup", str(warmup),
                                    "--Output", os.path.join(self.td.name, "o")] + ([] if converge else ["--NoOptimize"]))
            b = [r for r in recs if r["phase"] == "bench"][0]
            return float(b["seconds"]), recs
        ora = self.oracle_problem(threads)
        k = self.n_pc
        for _ in range(warmup):
            ora.compute_mix_llks([0.01] * k, [0.01] * k, 0.03)
        t0 = time.perf_counter()
        for i in range(evals):
            ora.compute_mix_llks([0.01 + 1e-6 * i] + [0.01] * (k - 1), [0.01] * k, 0.03)
        return time.perf_counter() - t0, None


def converge_reference(recs, threads):
    o = [r for r in recs if r["phase"] == "optimize"][0]
    load = [r for r in recs if r["phase"] == "load"][0]
    return {"optimize_s": o["optimize_s"], "evals": o["evals"], "threads": threads, "alpha": o["alpha"],
            "pcs": o["pc_contam"] + o["pc_intended"], "read_panel_s": load["panel_s"], "read_pileup_s": load["pileup_s"]}


def converge_ours(prefix, pile, td, n_pc):
    """Wall-clock to converged alpha through the product CLI (C++ host + GPU engine)."""
    import re
    from verifybamid_b200 import host
    out = os.path.join(td, "gpu")
    t0 = time.perf_counter()
    cp = subprocess.run([host.CLI_PATH, "--SVDPrefix", prefix, "--PileupFile", pile, "--Reference", "x", "--NumPC",
                         str(n_pc), "--Output", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    wall = time.perf_counter() - t0
    phase = dict(re.findall(r"Finished phase: (.*?)  \[([0-9.]+) seconds\]", cp.stderr))
    m = re.search(r"Likelihood evaluations: (\d+), ([0-9.]+) ms inside the GPU engine", cp.stderr)
    rows = [l.split("\t") for l in open(out + ".Ancestry").read().splitlines()[1:]]
    sm = open(out + ".selfSM").read().splitlines()[1].split("\t")
    return {"optimize_s": float(phase.get("Optimize likelihood", "nan")), "evals": int(m.group(1)),
            "engine_ms": float(m.group(2)), "flatten_upload_s": float(phase.get("  Flatten pileup into HBM", phase.get("Flatten pileup into HBM", "nan"))),
            "read_panel_s": float(phase.get("Load SVD reference data", "nan")), "read_pileup_s": float(phase.get("Read pileup", "nan")),
            "process_wall_s": wall, "alpha": float(sm[6]), "pcs": [float(r[1]) for r in rows] + [float(r[2]) for r in rows],
            "note": "alpha/PCs as printed (6 significant digits)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = make_workload(args.config)
    cfg = config_dict(args.config, sample, args.gpus)
    threads = os.cpu_count() or 1
    reads = cfg["reads_per_eval"]
    s_ref = min(REF_EVALS_PER_STEP, cfg["evals_per_step"])
    with CpuReference(sample) as ref:
        secs, _ = ref.time_evals(args.steps * s_ref, args.warmup * s_ref, threads)
        kind = ref.kind
    step_s = secs / args.steps
    value = reads * s_ref / step_s
    cpu = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": "every step evaluates a bounded sample of %d of the step's %d evaluations (all %d reads each), "
                     "%d OpenMP threads; %.3f ms per evaluation" % (s_ref, cfg["evals_per_step"], reads, threads,
                                                                    step_s / s_ref * 1e3)}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": cfg, "evals_timed_per_step": s_ref,
                      "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cpu_baseline_leg(sample, n_pc: int, want_converge: bool):
    """The reference CPU path at 1, 4 (the CLI's default --NumThread) and all threads; about 10-20 s of CPU work."""
    all_threads = os.cpu_count() or 1
    by_threads, conv = {}, None
    reads = sample.problem.used_counts()[1]
    with CpuReference(sample) as ref:
        for threads, n_eval in ((1, 12), (4, 40), (all_threads, 100)):
            if str(threads) in by_threads:
                continue
            converge = want_converge and threads == all_threads and ref.kind == "reference"
            secs, recs = ref.time_evals(n_eval, 3, threads, converge=converge)
            by_threads[str(threads)] = {"value": reads / (secs / n_eval), "ms_per_eval": secs / n_eval * 1e3, "evals": n_eval}
            if converge:
                conv = {"reference": converge_reference(recs, threads),
                        "ours": converge_ours(ref.prefix, ref.pile, ref.td.name, n_pc)}
                conv["abs_diff_alpha"] = abs(conv["ours"]["alpha"] - conv["reference"]["alpha"])
                conv["abs_diff_pc_max"] = max(abs(a - b) for a, b in zip(conv["ours"]["pcs"], conv["reference"]["pcs"]))
        kind = ref.kind
    top = by_threads[str(all_threads)]
    return {"value": top["value"], "unit": UNIT, "cores": all_threads, "kind": kind,
            "sample": "%d full evaluations of the same workload (%d reads each) on all %d threads (%.2f ms per evaluation); "
                      "12 at 1 thread, 40 at 4 threads (the reference CLI's default)" % (top["evals"], reads, all_threads, top["ms_per_eval"]),
            "by_threads": by_threads, "wall_clock_to_converged_alpha": conv}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import verifybamid_b200 as vb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the LLK engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # a non-default stream shared by torch (events, NCCL) and every engine context
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    panel_name, k, depth, n_samples, _, _ = CONFIGS[args.config]
    cohort = n_samples > 1
    sample = make_workload(args.config)
    cfg = config_dict(args.config, sample, world)
    p = sample.problem
    reads_per_eval, markers_used = cfg["reads_per_eval"], cfg["markers_used"]
    per_step = cfg["evals_per_step"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident samples ----------------------------------------------------------------------------------
    if cohort:
        # configs[3]: sample s lives on GPU s % world; a step evaluates every sample once, no collective
        mine = [s for s in range(n_samples) if s % world == rank]
        probs = [p if s == 0 else make_workload(args.config, seed=SEED + s).problem for s in mine]
        engines = [vb.LLKEngine(q, device=local, stream=stream.cuda_stream, batched=True) for q in probs]
        reads_step_total = None   # summed over ranks below
        launch_list = engines
        info = engines[0].info()
        copies = len(engines)
        reads_mine = float(sum(q.used_counts()[1] for q in probs))
        t = torch.tensor([reads_mine], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        reads_step_total = float(t.item())
    else:
        # enough resident copies of this rank's shard to exceed L2 -> every evaluation streams from HBM
        # (N > 1: a rank only ever evaluates its shard in batches -> the batched layout, include/vb2_llk.h)
        mk = lambda: vb.LLKEngine(p, device=local, shard_rank=rank, shard_count=world, stream=stream.cuda_stream, batched=world > 1)
        probe = mk()
        info = probe.info()
        copies = min(256, max(2, int(np.ceil(2.0 * L2_BYTES / max(1, info["device_bytes"])))))
        if world > 1:                               # shards differ slightly in size: every rank must agree
            c = torch.tensor([copies], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.MAX)
            copies = int(c.item())
        engines = [probe] + [mk() for _ in range(copies - 1)]
        launch_list = (engines * (per_step // copies + 1))[:per_step]   # evaluation i of a step runs on copy i % copies
        reads_step_total = float(reads_per_eval) * per_step

    n_jobs = len(launch_list)
    start_pc = np.full(k, 0.01)
    pcs = np.tile(start_pc, (n_jobs, 1)); als = np.full(n_jobs, 0.03)
    ctx_arr = vb.context_array(launch_list)
    d_step = [torch.zeros(n_jobs, dtype=torch.float64, device=dev) for _ in range(2)]
    collective = world > 1 and not cohort
    peer = None
    if collective and args.collective == "peer":
        def exchange(handle: bytes):
            mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
            every = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(every, mine)
            torch.cuda.synchronize()
            return [e.cpu().numpy().tobytes() for e in every]
        peer = vb.PeerReduce(local, rank, world, exchange)

    def shard_step(params_a, params_b, alphas, out):
        """This rank's shard of one step + the sum over the shards, left in `out` (device)."""
        if peer is not None:
            peer.eval_many_device(launch_list, params_a, params_b, alphas, out.data_ptr(), ctx_arr)
            return None
        vb.eval_many_device(launch_list, params_a, params_b, alphas, out.data_ptr(), ctx_arr)
        return dist.all_reduce(out, op=dist.ReduceOp.SUM, async_op=True)

    def keep_busy(seconds: float):
        # rank-local launches only (a time-based loop must not contain collectives)
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end:
            vb.time_device_many(launch_list, 0, 2, start_pc, start_pc, 0.03)

    # ---- value: device-timed, EXACTLY K steps; a step = one launch over the step's evaluations -----------------------
    with ClockSampler(local) as clocks:
        keep_busy(0.3)                      # let nvidia-smi attach before the timed region
        barrier()
        if not collective:
            # steps are issued back to back from C (vb2_llk_time_device_many: `warmup` untimed launches, then K timed
            # ones bracketed by CUDA events on the launching stream)
            barrier()
            dev_ms = vb.time_device_many(launch_list, args.warmup, args.steps, start_pc, start_pc, 0.03)
        else:
            # every step: this rank's marker shard for the step's parameter sets, then ONE NCCL all-reduce of the partial
            # sums (two result buffers: the all-reduce of one step runs on NCCL's stream while the next step's kernel
            # runs on ours)
            works = [None, None]
            turn = [0]

            def one_step():
                b = turn[0] = turn[0] ^ 1
                if works[b] is not None:
                    works[b].wait()          # (stream-level: our stream waits for that buffer's previous all-reduce)
                works[b] = shard_step(pcs, pcs, als, d_step[b])

            def drain():
                for w in works:
                    if w is not None:
                        w.wait()
            for _ in range(args.warmup):
                one_step()
            drain()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(args.steps):
                one_step()
            drain()
            ev1.record(stream)
            barrier()
            dev_ms = ev0.elapsed_time(ev1)
        plan = engines[0].batch_plan(n_jobs)   # which many-evaluations kernel, how many launches of it per step
        launches = (plan["kernel_launches"] + 1 + (1 if peer is not None else 0)) * args.steps   # ... + llk_reduce_kernel (+ llk_gather_kernel)
        keep_busy(0.5)                      # clocks under the same load, for the sampler
        barrier()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = reads_step_total / (ms_per_step * 1e-3)

    # ---- parity inside the run: one step's results (after the all-reduce) against the CPU oracle -------------------
    par_pcs = np.tile(np.asarray(sample.pc_contam, dtype=np.float64)[:k], (n_jobs, 1))
    par_pc2 = np.tile(np.asarray(sample.pc_intended, dtype=np.float64)[:k], (n_jobs, 1))
    par_al = np.full(n_jobs, ALPHA)
    par_pcs[:, 0] += 1e-3 * (np.arange(n_jobs) % 7)          # seven different points over the step
    if collective:
        w = shard_step(par_pcs, par_pc2, par_al, d_step[0])
        if w is not None:
            w.wait()
    else:
        vb.eval_many_device(launch_list, par_pcs, par_pc2, par_al, d_step[0].data_ptr(), ctx_arr)
    got = d_step[0].cpu().numpy()
    parity = None
    if rank == 0:
        from oracle import vb2_oracle as vo  # the checker, never the thing measured
        probs0 = [p] if not cohort else probs[:2]
        worst = 0.0
        for si, q in enumerate(probs0):
            ora = vo.Problem(q.ud, q.means, q.base_info_index, q.alt_base, q.info_offset, q.bases, q.quals, None,
                             q.sanity_disabled, q.avg_depth, q.sd_depth, min(16, os.cpu_count() or 1))
            if cohort:
                want = ora.compute_mix_llks(list(par_pcs[si]), list(par_pc2[si]), float(par_al[si]))
                worst = max(worst, abs(got[si] - want) / abs(want))
            else:   # every evaluation of the step against the oracle's value for its parameter set
                wants = [ora.compute_mix_llks(list(par_pcs[j]), list(par_pc2[j]), float(par_al[j])) for j in range(7)]
                for j in range(n_jobs):
                    worst = max(worst, abs(got[j] - wants[j % 7]) / abs(wants[j % 7]))
        parity = {"batched_rel": worst, "tolerance": PARITY_TOL,
                  "checked": "all %d evaluations of one step (7 distinct parameter sets) vs the CPU oracle" % n_jobs
                             if not cohort else "first %d samples of the step vs the CPU oracle" % len(probs0)}
        if not worst <= PARITY_TOL:
            raise SystemExit("bench.py: parity check failed: %r" % parity)

    # ---- roofline of the dominant kernel: the launch timed above ----------------------------------------------
    peak, peak_src = measured_peak_gbs()
    if cohort:
        alg_launch = float(sum(e.info()["algorithmic_bytes"] for e in engines))
        dev_launch = float(sum(e.info()["device_bytes"] for e in engines))
    else:
        alg_launch = float(info["algorithmic_bytes"]) * n_jobs   # this rank's shard x the launch's evaluations
        dev_launch = float(info["device_bytes"]) * n_jobs
    launch_us = ms_per_step * 1e3
    achieved = alg_launch / (launch_us * 1e-6) / 1e9
    one_ms = None
    if not cohort and world == 1:
        n1 = 500
        one_ms = vb.time_device(engines, 20, n1, start_pc, start_pc, 0.03) / n1
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture of the kernel, per evaluation of the headline
    # workload at N=1 (profiles/r02_llk_flow_kernel_digest.txt: 849,343,744 + 4,357,632 bytes for a launch of 120 evaluations;
    # profiles/r02_llk_stream_kernel_digest.txt: 14,846,177,000 + 15,367,424 bytes for 2,048): the stored image, no re-reads
    # (algorithmic: 7.573 MB per evaluation)
    per_eval_traffic = {"llk_flow_kernel": 7114178.0, "llk_stream_kernel": 7256614.0}[plan["kernel"]]
    traffic = per_eval_traffic * n_jobs if (args.config == "100k30x" and world == 1) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_note": "bytes per step of %d evaluations, from the committed ncu capture of this kernel (%.2f MB per "
                                "evaluation = the stored image, no re-reads)" % (n_jobs, per_eval_traffic / 1e6) if traffic else None,
                "peak_source": peak_src, "kernel": plan["kernel"], "evaluations_per_launch": n_jobs,
                "kernel_launches_per_step": plan["kernel_launches"], "evaluations_per_kernel_launch": plan["jobs_per_launch"],
                "launch_us": launch_us, "us_per_evaluation": launch_us / n_jobs * (1 if not cohort else 1),
                "us_per_evaluation_one_launch_each": one_ms * 1e3 if one_ms else None,
                "algorithmic_bytes_per_launch": alg_launch, "device_bytes_per_launch": dev_launch,
                "includes_allreduce": collective,
                "note": "bound by instruction issue, not by HBM: an FP64 warp-instruction holds a B200 sub-partition's issue slot for "
                        "two cycles (tools/microbench_mix.cu), so an evaluation costs >= 2*F + G cycles per sub-partition with "
                        "F = 1,939 fp64 and G ~ 2,280 other warp-instructions (10 fp64 per streamed read + ~90 per 32-marker slice): "
                        "3.1 us for this instruction mix, 1.97 us for the fp64 instructions alone (DESIGN.md section 4)"}

    # ---- e2e: the public C-ABI call with HOST buffers, host<->device traffic inside the timed region ----------------
    e2e_extra, r_last = {}, None
    if not cohort and world == 1:
        # the call as the estimator makes it: DEPENDENT evaluations of ONE sample -- a step is a run of simplex searches
        # (vb2_llk_minimize: host start point + model in, AmoebaMinimizer::Minimize on the device inside an evaluation
        # session, result out) that adds up to at least EVALS_PER_STEP evaluations; the metric counts the evaluations
        # the engine reports
        n_dep = args.steps * per_step
        n_warm = args.warmup * per_step // 8 + 16
        nm_start = list(start_pc) + list(start_pc) + [float(np.log(0.03 / 0.97))]
        nm_model = (list(range(k)), list(range(k, 2 * k)), 2 * k)
        searches_per_step, dev_evals, min_bytes, r_last = None, 0, (0, 0), None
        if not args.no_session:
            try:
                engines[0].session_begin()
            except vb.VB2Error:            # the sample does not fit on chip (deep coverage): one launch per evaluation
                args.no_session = True
        if not args.no_session and 2 * k + 1 <= vb.VB2_MIN_MAX_DIM:
            r0 = engines[0].minimize(nm_start, *nm_model, ftol=1e-8)           # (also the warm-up)
            searches_per_step = max(1, -(-per_step // max(1, r0["evals"])))
            for _ in range(max(0, args.warmup - 1)):
                engines[0].minimize(nm_start, *nm_model, ftol=1e-8)
            t0 = time.perf_counter()
            for i in range(args.steps * searches_per_step):
                nm_start[0] = 0.01 + 1e-7 * (i % 1000)                         # a different starting point every search
                r_last = engines[0].minimize(nm_start, *nm_model, ftol=1e-8)
                dev_evals += r_last["evals"]
            min_s = time.perf_counter() - t0
            min_bytes = (searches_per_step * (280 + 16 * (13 + 2 * k)), searches_per_step * 208)
        # the same dependent evaluations driven from the host, one vb2_llk_eval per evaluation (doorbell in, mailbox out)
        if args.no_session:   # (profiler runs: ncu serialises launches, a resident kernel would wait for a doorbell
            e2e_s, last = vb.time_host(engines, n_warm, n_dep, start_pc, start_pc, 0.03)  # that cannot ring)
        else:
            e2e_s, last = vb.time_host(engines[:1], n_warm, n_dep, start_pc, start_pc, 0.03)
            engines[0].session_end()
        last_pc = start_pc.copy(); last_pc[0] = 0.01 + 1e-7 * ((n_dep - 1 + n_warm) % 1000)
        e2e_point = (last_pc, start_pc, 0.03)
        # one launch per evaluation (no session), rotating through the resident copies (HBM-cold every evaluation)
        n_cold = min(n_dep, 4000)
        cold_s, _ = vb.time_host(engines, 50, n_cold, start_pc, start_pc, 0.03)
        # the batched public call: one vb2_llk_eval_many per step with host arrays in and out
        t0 = time.perf_counter()
        nb = max(2, min(args.steps, 10))
        for _ in range(nb):
            vb.eval_many(launch_list, pcs, pcs, als)
        batched_s = (time.perf_counter() - t0) / nb
        t0 = time.perf_counter()
        for i in range(200):
            engines[i % copies].compute_mix_llks(start_pc, start_pc, 0.03)
        e2e_extra = {"us_per_evaluation_host_driven_session": e2e_s / n_dep * 1e6,
                     "us_per_evaluation_one_launch_per_evaluation": cold_s / n_cold * 1e6,
                     "us_per_evaluation_batched_public_call": batched_s / n_jobs * 1e6,
                     "python_binding_us_per_evaluation": (time.perf_counter() - t0) / 200 * 1e6}
        h2d, d2h = (2 * k + 1) * 8 * per_step, 8 * per_step
        caller = ("C loop over vb2_llk_eval (host buffers), one launch per evaluation (--no-session)" if args.no_session else
                  "C loop over vb2_llk_eval (host buffers): %d dependent evaluations per step inside an evaluation session "
                  "(resident kernel, sample in shared memory, host-mapped doorbell/mailbox)" % per_step)
        e2e_evals_per_step = per_step
        if searches_per_step is not None:   # the headline: the searches on the device
            e2e_s = min_s
            e2e_evals_per_step = dev_evals / args.steps
            h2d, d2h = min_bytes
            caller = ("python: %d x vb2_llk_minimize per step (host start point + model in, Nelder-Mead next to the kernel inside "
                      "an evaluation session, result out): %.0f dependent evaluations per step" % (searches_per_step, e2e_evals_per_step))
            e2e_extra["searches_per_step"] = searches_per_step
            e2e_extra["evaluations_per_search"] = r0["evals"]
    else:
        # marker shards (or a cohort) go through the batched public call: per step, host parameter arrays in
        # (vb2_llk_eval_many_device stages them), this rank's launch, all-reduce (shards only), results back to pinned
        # host memory
        host_out = torch.zeros(n_jobs, dtype=torch.float64).pin_memory()
        host_pcs = pcs.copy()

        def e2e_step(i: int) -> float:
            host_pcs[:, 0] = 0.01 + 1e-7 * (i % 1000)
            if collective:
                w = shard_step(host_pcs, pcs, als, d_step[0])
                if w is not None:
                    w.wait()
            else:
                vb.eval_many_device(launch_list, host_pcs, pcs, als, d_step[0].data_ptr(), ctx_arr)
            host_out.copy_(d_step[0], non_blocking=False)
            return float(host_out[n_jobs - 1])
        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        last = 0.0
        for i in range(args.steps):
            last = e2e_step(i)
        barrier()
        e2e_s = time.perf_counter() - t0
        last_pc = start_pc.copy(); last_pc[0] = 0.01 + 1e-7 * ((args.steps - 1) % 1000)
        e2e_point = (last_pc, start_pc, 0.03)
        h2d, d2h = n_jobs * (2 * k + 1) * 8, n_jobs * 8
        caller = "python: vb2_llk_eval_many_device over %d host parameter sets per step%s + D2H of the results" % (
            n_jobs, (" (marker shard) + " + ("peer stores over NVLink" if peer is not None else "NCCL all-reduce")) if collective
            else " (this GPU's samples)")
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_reads_step = reads_step_total if (cohort or world > 1) else float(reads_per_eval) * e2e_evals_per_step
    e2e_n = n_jobs if (cohort or world > 1) else e2e_evals_per_step
    e2e = {"value": e2e_reads_step / (e2e_s / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "us_per_evaluation": e2e_s / args.steps / e2e_n * 1e6, "ms_per_step": e2e_s / args.steps * 1e3,
           "evaluations_per_step": e2e_n, "last_llk": last, "caller": caller}
    e2e.update(e2e_extra)
    if rank == 0:   # the e2e result against the oracle too (cohort: the last sample of rank 0)
        from oracle import vb2_oracle as vo
        q = p if not cohort else probs[-1]
        ora = vo.Problem(q.ud, q.means, q.base_info_index, q.alt_base, q.info_offset, q.bases, q.quals, None,
                         q.sanity_disabled, q.avg_depth, q.sd_depth, min(16, os.cpu_count() or 1))
        want = ora.compute_mix_llks(list(e2e_point[0]), list(e2e_point[1]), e2e_point[2])
        parity["e2e_rel"] = abs(last - want) / abs(want)
        if not parity["e2e_rel"] <= PARITY_TOL:
            raise SystemExit("bench.py: e2e parity check failed: got %.17g want %.17g" % (last, want))
        if not cohort and world == 1 and r_last is not None:   # the last search on the device: its minimum at its best point
            v = r_last["point"]
            a = float(np.exp(v[2 * k])); a = a / (1.0 + a)
            want = -ora.compute_mix_llks(list(v[:k]), list(v[k:2 * k]), a)
            parity["search_fmin_rel"] = abs(r_last["fmin"] - want) / abs(want)
            parity["search_alpha"] = a
            if not (parity["search_fmin_rel"] <= PARITY_TOL and r_last["converged"]):
                raise SystemExit("bench.py: device search parity check failed: %r" % (r_last,))

    # ---- cpu baseline beside it (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(sample, k, want_converge=(args.config != "batch64"))

    if peer is not None:
        barrier()
        peer.close()
    for e in engines:
        e.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "us_per_evaluation": ms_per_step * 1e3 / n_jobs if not cohort else None,
                "cache": "a step rotates through %d resident copies of the sample (%.0f MB > 126 MB L2): every evaluation streams "
                         "from HBM" % (copies, copies * info["device_bytes"] / 1e6) if not cohort else
                         "%d resident samples on this GPU (%.0f MB > 126 MB L2)" % (copies, dev_launch / 1e6),
                "precision": "fp32 UD/mu in HBM, fp64 arithmetic",
                "collective": (None if not collective else "peer stores over NVLink from llk_reduce_kernel + llk_gather_kernel "
                               "(%d bytes pushed per rank and step)" % (8 * n_jobs * world) if peer is not None else
                               "NCCL all-reduce of %d doubles per step, overlapped with the next step's kernel" % n_jobs),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "parity": parity,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="100k30x", choices=sorted(CONFIGS))
    ap.add_argument("--collective", default="nccl", choices=["nccl", "peer"],
                    help="N>1: how the shard sums meet -- an NCCL all-reduce behind the kernel, or peer stores over NVLink from "
                         "the reduce kernel itself (vb2_peer_*)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-session", action="store_true", help="e2e with one launch per evaluation (for runs under ncu)")
    args = ap.parse_args()
    if args.steps is None:            # defaults that finish within minutes on either arm
        args.steps = 20
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
