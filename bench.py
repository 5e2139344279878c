#!/usr/bin/env python
"""bench.py -- marker-reads/sec per LLK evaluation on BASELINE.json configs[1]
(synthetic pileup, 1000g.phase3.100k.b37 panel, 100k markers x 30x, NumPC=2, alpha=0.02).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one evaluation of the contamination log-likelihood (one call of the reference's
ComputeMixLLKs) over the whole sample.  With N GPUs the markers are sharded across ranks
(32-marker slices, round-robin) and a step is: every rank evaluates its shard, then ONE NCCL
allreduce of the scalar partial -- strong scaling of a fixed sample.

Keys of the JSON line (rank 0):
  value     marker-reads/s with everything resident in HBM, device-timed (CUDA events on the launching
            stream) over K back-to-back steps that rotate through enough resident copies of the sample
            to exceed L2, so every step streams from HBM;
  e2e       the same metric through the public call (vb2_llk_eval) with HOST parameter buffers, every step moving
            its inputs (2k+1 doubles) to the device and its scalar result back, inside an evaluation session
            (resident kernel, sample in shared memory) -- the way the simplex search calls it; the figure for one
            launch per evaluation is reported beside it;
  roofline  algorithmic bytes (2*R + 4*(k+2)*M') / measured kernel time vs measured HBM peak;
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, else the C port) on this host.
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
PANEL = "1000g.phase3.100k.b37"
N_PC, DEPTH, ALPHA, SEED = 2, 30.0, 0.02, 1
WORKLOAD = "synthetic pileup, %s panel, 100k markers x 30x, NumPC=2, alpha=0.02, sanity filter on (seed 1)" % PANEL
L2_BYTES = 126 * 1024 * 1024


def make_workload():
    from verifybamid_b200 import panels, synth
    panel = panels.load_bundled(PANEL)
    return synth.make_sample(panel, n_pc=N_PC, depth=DEPTH, alpha=ALPHA, seed=SEED, sanity_check=True)


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
def cpu_reference_run(sample, evals: int, warmup: int, threads: int, converge: bool = False):
    """Time `evals` evaluations of the whole workload on the CPU.  Returns (seconds, kind, reads_used, conv)
    where conv (when asked for) compares the wall-clock to converged alpha of the C++ GPU CLI and of the
    reference binary on the same panel + pileup text files."""
    from oracle import vb2_oracle as vo  # checker / baseline only -- never on the product path
    p = sample.problem
    conv = None
    if vo.ref_available():
        from verifybamid_b200 import panels
        with tempfile.TemporaryDirectory() as td:
            prefix = panels.write_text_panel(sample.panel, os.path.join(td, "panel"))
            pile = sample.write_pileup(os.path.join(td, "sample.pileup"))
            recs = vo.run_ref(["--SVDPrefix", prefix, "--PileupFile", pile, "--NumPC", str(N_PC), "--NumThread",
                               str(threads), "--BenchEvals", str(evals), "--BenchWarmup", str(warmup),
                               "--Output", os.path.join(td, "o")] + ([] if converge else ["--NoOptimize"]))
            if converge:
                conv = {"reference": converge_reference(recs, threads), "ours": converge_ours(prefix, pile, td)}
                conv["abs_diff_alpha"] = abs(conv["ours"]["alpha"] - conv["reference"]["alpha"])
                conv["abs_diff_pc_max"] = max(abs(a - b) for a, b in zip(conv["ours"]["pcs"], conv["reference"]["pcs"]))
        b = [r for r in recs if r["phase"] == "bench"][0]
        return float(b["seconds"]), "reference", int(b["reads_used"]), conv
    ora = vo.Problem(p.ud, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, None,
                     p.sanity_disabled, p.avg_depth, p.sd_depth, threads)
    for _ in range(warmup):
        ora.compute_mix_llks([0.01] * N_PC, [0.01] * N_PC, 0.03)
    t0 = time.perf_counter()
    for i in range(evals):
        ora.compute_mix_llks([0.01 + 1e-6 * i] + [0.01] * (N_PC - 1), [0.01] * N_PC, 0.03)
    return time.perf_counter() - t0, "port", ora.used_counts()[1], conv


def converge_reference(recs, threads):
    o = [r for r in recs if r["phase"] == "optimize"][0]
    load = [r for r in recs if r["phase"] == "load"][0]
    return {"optimize_s": o["optimize_s"], "evals": o["evals"], "threads": threads, "alpha": o["alpha"],
            "pcs": o["pc_contam"] + o["pc_intended"], "read_panel_s": load["panel_s"], "read_pileup_s": load["pileup_s"]}


def converge_ours(prefix, pile, td):
    """Wall-clock to converged alpha through the product CLI (C++ host + GPU engine)."""
    import re
    from verifybamid_b200 import host
    out = os.path.join(td, "gpu")
    t0 = time.perf_counter()
    cp = subprocess.run([host.CLI_PATH, "--SVDPrefix", prefix, "--PileupFile", pile, "--Reference", "x", "--NumPC",
                         str(N_PC), "--Output", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
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
    sample = make_workload()
    threads = os.cpu_count() or 1
    secs, kind, reads, _ = cpu_reference_run(sample, args.steps, args.warmup, threads)
    ms = secs / args.steps * 1e3
    value = reads / (secs / args.steps)
    cpu = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": "%d full evaluations of the workload (all %d reads each), %d threads" % (args.steps, reads, threads)}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "reads_per_step": reads},
                      "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import verifybamid_b200 as vb
    from verifybamid_b200.distributed import allreduce_partials

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

    sample = make_workload()
    p = sample.problem
    k = p.n_pc
    # enough resident copies of this rank's shard to exceed L2 -> every step streams from HBM
    # (N > 1: every rank only ever evaluates its shard in batches -> the batched layout, include/vb2_llk.h)
    probe = vb.LLKEngine(p, device=local, shard_rank=rank, shard_count=world, stream=stream.cuda_stream, batched=world > 1)
    info = probe.info()
    copies = max(2, int(np.ceil(2.0 * L2_BYTES / max(1, info["device_bytes"]))))
    copies = min(copies, 256)
    if world > 1:                               # shards differ slightly in size: every rank must agree
        c = torch.tensor([copies], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        copies = int(c.item())
    engines = [probe] + [vb.LLKEngine(p, device=local, shard_rank=rank, shard_count=world, stream=stream.cuda_stream, batched=world > 1)
                         for _ in range(copies - 1)]
    reads_total = p.used_counts()[1]           # whole sample, all shards
    markers_total = p.used_counts()[0]
    d_out = torch.zeros(1, dtype=torch.float64, device=dev)
    pc_a = np.full((1, k), 0.01); pc_b = np.full((1, k), 0.01); al = np.array([0.03])

    def step(i: int):
        # rank-local kernel on this rank's marker shard, then ONE allreduce of the scalar partial
        pc_a[0, 0] = 0.01 + 1e-7 * (i % 1000)
        engines[i % copies].eval_batch_device(pc_a, pc_b, al, d_out.data_ptr())
        allreduce_partials(d_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # A launch makes ROTATIONS passes over the resident copies (a context may appear several times in a launch):
    # the ramp and the tail of the persistent kernel are shared by more steps.
    ROTATIONS = 8
    per_launch = min(copies * ROTATIONS, 2048, max(copies, args.steps // 4))  # (at least four launches to pipeline)
    launch_list = (engines * ROTATIONS)[:per_launch]

    start_pc = np.full(k, 0.01)

    def keep_busy(seconds: float):
        # rank-local launches only (a time-based loop must not contain collectives)
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end:
            vb.time_device_many(launch_list, 0, 3, start_pc, start_pc, 0.03)

    # ---- value: device-timed, EXACTLY K steps ---------------------------------------------------
    # N=1: steps are issued from C, `copies` steps per launch (vb2_llk_eval_many's kernel: step i
    # evaluates resident copy i % copies, so every step streams its sample from HBM and the launch
    # cost is shared by the steps of a launch).  N>1: each step is kernel + NCCL allreduce of the
    # scalar, one launch per step, issued from Python.
    def timed_steps(n_steps: int, warm: int) -> float:
        full, rem = divmod(n_steps, per_launch)
        ms = 0.0
        if full:
            ms += vb.time_device_many(launch_list, max(1, warm // per_launch), full, start_pc, start_pc, 0.03)
        if rem:
            ms += vb.time_device_many(launch_list[:rem], 1, 1, start_pc, start_pc, 0.03)
        return ms
    with ClockSampler(local) as clocks:
        keep_busy(0.3)                      # let nvidia-smi attach before the timed region
        barrier()
        if world == 1:
            dev_ms = timed_steps(args.steps, args.warmup)
            launches = 2 * (args.steps // per_launch + (1 if args.steps % per_launch else 0))  # stream kernel + reduce kernel
        else:
            # `per_launch` steps per launch on every rank (its marker shard of each resident copy), then ONE
            # NCCL allreduce of their partial sums
            # (two result buffers: the allreduce of one launch's partial sums runs on NCCL's stream while the next
            # launch's kernel runs on ours)
            d_many = [torch.zeros(per_launch, dtype=torch.float64, device=dev) for _ in range(2)]
            works = [None, None]
            pcs = np.tile(start_pc, (per_launch, 1)); als = np.full(per_launch, 0.03)
            turn = [0]
            ctx_arr = vb.context_array(launch_list)

            def launch_steps(n: int):
                b = turn[0] = turn[0] ^ 1
                if works[b] is not None:
                    works[b].wait()          # (stream-level: our stream waits for that buffer's previous allreduce)
                vb.eval_many_device(launch_list[:n], pcs[:n], pcs[:n], als[:n], d_many[b].data_ptr(), ctx_arr)
                works[b] = dist.all_reduce(d_many[b][:n], op=dist.ReduceOp.SUM, async_op=True)

            def drain():
                for w in works:
                    if w is not None:
                        w.wait()
            full, rem = divmod(args.steps, per_launch)
            for _ in range(max(1, args.warmup // per_launch)):
                launch_steps(per_launch)
            drain()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(full):
                launch_steps(per_launch)
            if rem:
                launch_steps(rem)
            drain()
            ev1.record(stream)
            barrier()
            dev_ms = ev0.elapsed_time(ev1)
            launches = 2 * (full + (1 if rem else 0))  # stream kernel + reduce kernel per launch_steps()
        keep_busy(0.5)                      # clocks under the same load, for the sampler
        barrier()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = reads_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant (only) kernel -------------------------------------------------
    # N=1: the kernel of the timed region above (algorithmic bytes of one launch / its duration).
    # Also reported: the same kernel launched once per step (latency geometry), which on this part
    # cannot go below the ~4 us cost of any launch that contains a block-wide barrier.
    one_ms = vb.time_device(engines, args.warmup, min(args.steps, 2000), start_pc, start_pc, 0.03) / min(args.steps, 2000)
    peak, peak_src = measured_peak_gbs()
    alg_bytes = info["algorithmic_bytes"]       # this rank's shard, one evaluation
    kern_us = (dev_ms / args.steps) * 1e3       # per evaluation (N>1: includes the allreduce share)
    achieved = alg_bytes / (kern_us * 1e-6) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture of this kernel (296 evaluations in the
    # launch: 2,145,602,000 + 5,224,192 bytes; profiles/r01_llk_stream_kernel_ncu_details.txt), per evaluation:
    traffic_per_eval = 7266305 if world == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_per_eval * per_launch if traffic_per_eval else None,
                "traffic_note": "bytes per launch of %d evaluations, from the committed ncu capture (7.27 MB per evaluation = "
                                "the stored image, no re-reads; algorithmic 7.57 MB)" % per_launch, "peak_source": peak_src, "kernel": "llk_stream_kernel",
                "evaluations_per_launch": per_launch,
                "us_per_evaluation": kern_us, "us_per_evaluation_one_launch_each": one_ms * 1e3,
                "algorithmic_bytes_per_evaluation": alg_bytes, "device_bytes_per_evaluation": info["device_bytes"],
                "note": "co-bound by the FP64 pipe: 10 fp64 instructions per streamed read + ~40 per marker -> >= 2.2 us per "
                        "evaluation at 64 lanes/clk/SM (DESIGN.md section 4)"}

    # ---- e2e: the public C-ABI call with HOST buffers; every step moves the step's inputs (2k+1 doubles)
    # to the device and the scalar result back to the host ------------------------------------------------
    if world == 1:
        # (a) the call as the simplex search makes it: several hundred dependent evaluations of ONE sample inside an
        #     evaluation session (vb2_llk_session_begin: resident kernel, the sample in shared memory, host-mapped
        #     doorbell in, host mailbox out).  Every step still moves its 2k+1 doubles in and its scalar out.
        if args.no_session:   # (profiler runs: ncu serialises launches, a resident kernel would wait for a doorbell
            e2e_s, last = vb.time_host(engines, args.warmup, args.steps, start_pc, start_pc, 0.03)  # that cannot ring)
        else:
            engines[0].session_begin()
            e2e_s, last = vb.time_host(engines[:1], args.warmup, args.steps, start_pc, start_pc, 0.03)
            engines[0].session_end()
        # (b) one launch per evaluation (no session), rotating through the resident copies (HBM-cold every step)
        cold_s, last_cold = vb.time_host(engines, args.warmup, args.steps, start_pc, start_pc, 0.03)
        # the same call through the Python binding (interpreter + ctypes overhead included)
        t0 = time.perf_counter()
        for i in range(200):
            engines[i % copies].compute_mix_llks(start_pc, start_pc, 0.03)
        py_us = (time.perf_counter() - t0) / 200 * 1e6
    else:
        # Marker shards cannot shorten ONE dependent evaluation (a ~3 us kernel against a ~30 us collective), so the
        # sharded public call is the batched one: every call takes `copies` parameter sets from HOST arrays
        # (vb2_llk_eval_many_device stages them: 508 bytes per evaluation), evaluates this rank's shard of each,
        # all-reduces the partial sums and copies the `copies` results back to pinned host memory.
        host_out = torch.zeros(copies, dtype=torch.float64).pin_memory()
        d_e2e = torch.zeros(copies, dtype=torch.float64, device=dev)
        host_pcs = np.tile(start_pc, (copies, 1)); host_als = np.full(copies, 0.03)

        def e2e_batch(i: int, n: int) -> float:
            host_pcs[:, 0] = 0.01 + 1e-7 * (i % 1000)
            vb.eval_many_device(engines[:n], host_pcs[:n], host_pcs[:n], host_als[:n], d_e2e.data_ptr())
            allreduce_partials(d_e2e[:n])
            host_out[:n].copy_(d_e2e[:n], non_blocking=False)
            return float(host_out[n - 1])
        for i in range(max(1, args.warmup // copies)):
            e2e_batch(i, copies)
        barrier()
        t0 = time.perf_counter()
        last = 0.0
        full, rem = divmod(args.steps, copies)
        for i in range(full):
            last = e2e_batch(i, copies)
        if rem:
            last = e2e_batch(full, rem)
        barrier()
        e2e_s = time.perf_counter() - t0
        py_us = None
        cold_s = None
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": reads_total / (e2e_s / args.steps), "unit": UNIT, "h2d_bytes_per_step": (2 * k + 1) * 8 if world == 1 else 508,
           "d2h_bytes_per_step": 8, "us_per_step": e2e_s / args.steps * 1e6, "last_llk": last,
           "caller": ("C loop over vb2_llk_eval (host buffers), one launch per evaluation (--no-session)" if args.no_session else
                      "C loop over vb2_llk_eval (host buffers) inside an evaluation session: resident kernel, sample in "
                      "shared memory, host-mapped doorbell/mailbox") if world == 1 else
                     "python: vb2_llk_eval_many_device over %d host parameter sets per call (marker shard) + NCCL allreduce + "
                     "D2H of the results" % copies,
           "us_per_step_one_launch_per_evaluation": (cold_s / args.steps * 1e6) if cold_s else None,
           "python_binding_us_per_step": py_us}

    # ---- cpu baseline beside it (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_eval = 100
        secs, kind, reads, conv = cpu_reference_run(sample, n_eval, 3, threads, converge=True)
        cpu = {"value": reads / (secs / n_eval), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d full evaluations of the same workload (%d reads each), %d OpenMP threads; %.2f ms/eval"
                         % (n_eval, reads, threads, secs / n_eval * 1e3),
               "wall_clock_to_converged_alpha": conv}

    for e in engines:
        e.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "reads_per_step": reads_total, "markers_used": markers_total,
                           "n_pc": k, "parallelism": "marker shards x%d + 1 scalar allreduce/step" % world if world > 1
                           else "single GPU", "l2": "steps rotate through %d resident copies of the sample "
                           "(%.0f MB > 126 MB L2): every step streams from HBM" % (copies, copies * info["device_bytes"] / 1e6),
                           "steps_per_launch": per_launch,
                           "panel_dtype": "fp32 UD/mu in HBM, fp64 arithmetic"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-session", action="store_true", help="e2e with one launch per evaluation (for runs under ncu)")
    args = ap.parse_args()
    if args.steps is None:            # defaults that finish within minutes on either arm
        args.steps = 2000 if args.impl == "ours" else 200
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
