"""Stage 1 over several GPUs: the IMAGE dataset shards (rank r tallies a contiguous range), and — unlike the
describe path, where neurons are independent — the per-rank statistics have to be exchanged:

  * top-k lists        all-gather of (units, k) values + dataset indices, then the same "larger value, earlier
                       index" selection the tally kernel uses: identical to a single-GPU tally
  * quantile samples   exact regime (<= 8192 samples per unit in TOTAL): all-gather of the kept samples; the
                       estimator sorts them, so the result is identical to a single-GPU tally
  * histograms         beyond it: one all-reduce(SUM) of the (units, 65536) integer histograms (adds commute)
  * masks / images     each (unit, rank-in-top-k) slot is produced by exactly one rank: all-reduce(MAX) of the
                       uint8 result tensors

NCCL on GPUs; the host-side merge logic is exercised over gloo on CPU tensors (`tests/test_sharding.py`).
The reference has no multi-GPU path; this is new.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def world_and_rank() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def select_topk(values: torch.Tensor, ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(units, m) candidate values / dataset indices (index < 0 = empty slot) -> the k best per unit, ordered by
    larger value then earlier index, empty slots last (values -inf, ids -1)."""
    empty = ids < 0
    values = values.masked_fill(empty, float('-inf'))
    big = torch.iinfo(torch.long).max
    order = torch.argsort(ids.masked_fill(empty, big), dim=1, stable=True)  # by index ...
    values, ids = values.gather(1, order), ids.gather(1, order)
    order = torch.argsort(values, dim=1, descending=True, stable=True)       # ... then stably by value
    values, ids = values.gather(1, order)[:, :k], ids.gather(1, order)[:, :k]
    return values, ids


def _all_gather(tensor: torch.Tensor) -> List[torch.Tensor]:
    world, _ = world_and_rank()
    out = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(out, tensor.contiguous())
    return out


def merge_topk(values: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """This rank's (units, k) top-k -> the top-k over all ranks (same on every rank)."""
    world, _ = world_and_rank()
    if world == 1:
        return values, ids
    k = values.shape[1]
    return select_topk(torch.cat(_all_gather(values), dim=1), torch.cat(_all_gather(ids), dim=1), k)


def gather_samples(samples: torch.Tensor, count: int, capacity: int) -> Tuple[torch.Tensor, int]:
    """Exact regime: this rank's kept samples (units, capacity) with `count` valid columns -> all ranks' samples
    concatenated (units, capacity), total count. The caller guarantees total <= capacity."""
    world, _ = world_and_rank()
    if world == 1:
        return samples, count
    counts = [torch.zeros(1, dtype=torch.long, device=samples.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([count], dtype=torch.long, device=samples.device))
    counts = [int(c.item()) for c in counts]
    parts = _all_gather(samples)
    merged = torch.empty_like(samples)
    at = 0
    for part, n in zip(parts, counts):
        merged[:, at:at + n] = part[:, :n]
        at += n
    return merged, at


def total_count(count: int, device) -> int:
    world, _ = world_and_rank()
    if world == 1:
        return count
    t = torch.tensor([count], dtype=torch.long, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def max_over_ranks(value: int, device) -> int:
    world, _ = world_and_rank()
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.long, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def sum_histograms(hist: torch.Tensor) -> torch.Tensor:
    world, _ = world_and_rank()
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def max_bytes(tensor: torch.Tensor) -> torch.Tensor:
    """uint8 results where every slot is written by exactly one rank (zeros elsewhere) -> the union."""
    world, _ = world_and_rank()
    if world > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
    return tensor
