#!/usr/bin/env python
"""A/B of the two ways the marker shards' sums meet (run under torchrun, one rank per GPU):
NCCL all-reduce behind the kernel  vs  peer stores over NVLink from the reduce kernel + gather kernel (vb2_peer_*).
  (a) one step of EVALS evaluations, device-timed back to back;  (b) ONE dependent evaluation, host-timed to the scalar."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import verifybamid_b200 as vb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
s = bench.make_workload("100k30x")
p = s.problem
k = p.n_pc
engines = [vb.LLKEngine(p, device=local, shard_rank=rank, shard_count=world, stream=stream.cuda_stream, batched=True) for _ in range(8)]
EVALS = 2048
lst = (engines * (EVALS // 8 + 1))[:EVALS]
arr = vb.context_array(lst)
pcs = np.tile(np.full(k, 0.01), (EVALS, 1)); als = np.full(EVALS, 0.03)
out = [torch.zeros(EVALS, dtype=torch.float64, device=dev) for _ in range(2)]


def exchange(h):
    mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(dev)
    every = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(every, mine)
    torch.cuda.synchronize()
    return [e.cpu().numpy().tobytes() for e in every]


peer = vb.PeerReduce(local, rank, world, exchange)


def step_nccl(n, o):
    vb.eval_many_device(lst[:n], pcs[:n], pcs[:n], als[:n], o.data_ptr(), arr)
    dist.all_reduce(o[:n], op=dist.ReduceOp.SUM)


def step_peer(n, o):
    peer.eval_many_device(lst[:n], pcs[:n], pcs[:n], als[:n], o.data_ptr(), arr)


def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


res = {}
for name, fn in (("nccl", step_nccl), ("peer", step_peer)):
    for _ in range(5):
        fn(EVALS, out[0])
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(20):
        fn(EVALS, out[i & 1])
    e1.record(stream)
    sync()
    res[name + "_step_ms"] = e0.elapsed_time(e1) / 20
    res[name + "_values"] = out[1][:3].cpu().numpy().tolist()
    for _ in range(20):
        fn(1, out[0]); float(out[0][0])
    sync()
    t0 = time.perf_counter()
    for _ in range(200):
        fn(1, out[0]); v = float(out[0][0])
    res[name + "_single_eval_us"] = (time.perf_counter() - t0) / 200 * 1e6
    sync()
t = torch.tensor([res["nccl_step_ms"], res["peer_step_ms"], res["nccl_single_eval_us"], res["peer_single_eval_us"]], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    a = t.cpu().numpy()
    # (the peer path adds the shards in rank order on every rank; NCCL adds them in its ring/tree order: the sums may
    #  differ in the last bit from two ranks on)
    worst = max(abs(x - y) / abs(y) for x, y in zip(res["nccl_values"], res["peer_values"]))
    print("collective A/B, %d GPUs, %d evaluations per step: NCCL %.4f ms/step, peer %.4f ms/step | one dependent evaluation to the "
          "host scalar: NCCL %.1f us, peer %.1f us | NVLink bytes pushed per rank and step: %d | largest relative difference "
          "of the sums: %.1e (%r vs %r)"
          % (world, EVALS, a[0], a[1], a[2], a[3], 8 * EVALS * (world - 1), worst, res["nccl_values"][0], res["peer_values"][0]))
peer.close()
for e in engines:
    e.close()
dist.destroy_process_group()
