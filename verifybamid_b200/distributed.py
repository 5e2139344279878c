"""Marker-sharded evaluation over several GPUs, one process per GPU (torch.distributed).

The log-likelihood is a plain sum over markers (the reference reduces it with
`#pragma omp parallel for reduction(+:sumLLK)`, ContaminationEstimator.h:232-236), so rank r keeps shard r
of the 32-marker slices (slice s -> rank s % world, see include/vb2_llk.h) resident in its own HBM and an
evaluation is: every rank launches its shard, then ONE all-reduce of the scalar partial sums.  There is no
other data-path collective.  torch.distributed is plumbing only (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .engine import LLKEngine, VB2_PANEL_FP32
from .problem import PileupProblem


def allreduce_partials(partials: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the per-rank partial log-likelihoods in place (fp64) and return the tensor."""
    if partials.dtype != torch.float64:
        raise TypeError("partial log-likelihoods are fp64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


class ShardedLLK:
    """ComputeMixLLKs over `world` marker shards; every rank gets the full sum."""

    def __init__(self, problem: PileupProblem, device: int, rank: int, world: int, panel_dtype: int = VB2_PANEL_FP32,
                 stream: Optional[torch.cuda.Stream] = None, group: Optional[dist.ProcessGroup] = None):
        self.rank, self.world, self.group = rank, world, group
        self.dev = torch.device("cuda", device)
        # kernel, collective and the read-back of the sum must be ordered on ONE stream torch knows about
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.dev)
        self.engine = LLKEngine(problem, device=device, panel_dtype=panel_dtype, shard_rank=rank, shard_count=world,
                                stream=self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            self._out = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.n_pc = problem.n_pc

    def launch(self, pc_contam: Sequence[float], pc_intended: Sequence[float], alpha: float) -> torch.Tensor:
        """Asynchronous: kernel on this rank's shard + all-reduce; the sum stays in device memory."""
        a = np.ascontiguousarray(pc_contam, dtype=np.float64).reshape(1, self.n_pc)
        b = np.ascontiguousarray(pc_intended, dtype=np.float64).reshape(1, self.n_pc)
        with torch.cuda.stream(self.stream):
            self.engine.eval_batch_device(a, b, np.array([alpha]), self._out.data_ptr())
            return allreduce_partials(self._out, self.group)

    def compute_mix_llks(self, pc_contam: Sequence[float], pc_intended: Sequence[float], alpha: float) -> float:
        with torch.cuda.stream(self.stream):
            return float(self.launch(pc_contam, pc_intended, alpha).item())

    def close(self) -> None:
        self.engine.close()
