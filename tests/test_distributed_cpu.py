"""The N>1 path on CPU: two gloo ranks, each flattening ITS marker shard with the engine's own packer
(libvb2llk.so, host only) and evaluating it with the numpy restatement of the kernel; one all-reduce of the
scalar partials must reproduce the single-shard likelihood and the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import verifybamid_b200 as vb
from verifybamid_b200 import panels, synth
from verifybamid_b200.distributed import allreduce_partials
from helpers import emulate_packed_llk, to_oracle

POINT = ([0.02, -0.01], [-0.01, 0.027], 0.05)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        panel = panels.load_bundled("1000g.phase3.10k.b37")
        p = synth.make_sample(panel, n_pc=2, depth=12.0, alpha=0.05, seed=11, n_markers=1500).problem
        pk = vb.pack_host(p, shard_rank=rank, shard_count=world)          # this rank's shard only
        part = torch.tensor([emulate_packed_llk(pk, *POINT)], dtype=torch.float64)
        mine = float(part[0])
        total = float(allreduce_partials(part)[0])
        q.put((rank, pk["n_used"], pk["reads_used"], mine, total))
    finally:
        dist.destroy_process_group()


def test_two_rank_marker_shards_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    p = synth.make_sample(panel, n_pc=2, depth=12.0, alpha=0.05, seed=11, n_markers=1500).problem
    whole = vb.pack_host(p)
    assert sum(r[1] for r in res) == whole["n_used"] and sum(r[2] for r in res) == whole["reads_used"]
    assert res[0][4] == res[1][4]                                        # every rank holds the same sum
    assert abs(res[0][4] - (res[0][3] + res[1][3])) <= 1e-9
    want = to_oracle(p).compute_mix_llks(*POINT)
    assert abs(res[0][4] - want) <= 1e-10 * abs(want)
    assert abs(res[0][3]) > 0 and abs(res[1][3]) > 0                     # both shards carry work


def test_allreduce_is_identity_without_a_process_group():
    t = torch.tensor([-1.5], dtype=torch.float64)
    assert float(allreduce_partials(t)[0]) == -1.5
