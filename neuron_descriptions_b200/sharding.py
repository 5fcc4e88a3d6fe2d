"""Neuron sharding across the GPUs of one box.

Neurons are independent: rank r describes the contiguous range [r*ceil(N/G), (r+1)*ceil(N/G)) of the dataset
order with a full replica of the (~280 MB) weights; there is no exchange during compute. The one collective is an
all-gather of the final token ids (int64, padded to equal shard length) at the end — NCCL over NVLink on GPUs,
gloo in the CPU tests — after which every rank (rank 0 matters) detokenises in dataset order.
The reference has no multi-GPU path; this is new.
"""
import os
import pathlib
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; the last shards may be short or empty."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def init_distributed() -> Tuple[int, int, int]:
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group('gloo')
    return world, rank, local_rank


def barrier(device=None):
    """All ranks (no-op for one) + a device sync, so that a host clock read after it brackets GPU work."""
    if dist.is_initialized():
        dist.barrier()
    if device is not None and torch.cuda.is_available() and str(device).startswith('cuda'):
        torch.cuda.synchronize(device)


def finalize_distributed():
    if dist.is_initialized():
        dist.destroy_process_group()


def gather_token_rows(local: torch.Tensor, n_total: int, world: int, pad_value: int) -> torch.Tensor:
    """All-gather per-rank (n_local, T) int64 rows into (n_total, T) in dataset order."""
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    T = local.shape[1]
    padded = torch.full((per, T), pad_value, dtype=torch.long, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty(world * per, T, dtype=torch.long, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return out[:n_total]


class _Slice(torch.utils.data.Dataset):
    """Rows [lo, hi) of `base`, keeping the attributes `Decoder.predict` inspects (k, transforms)."""

    def __init__(self, base, lo, hi):
        self.base, self.lo, self.hi = base, lo, hi
        for name in ('k', 'transform_images', 'transform_masks'):
            if hasattr(base, name):
                setattr(self, name, getattr(base, name))

    def __len__(self):
        return self.hi - self.lo

    def __getitem__(self, index):
        return self.base[self.lo + index]



class _SliceU8(_Slice):
    def batch_u8(self, lo, hi, out=None):
        return self.base.batch_u8(self.lo + lo, self.lo + hi, out=out)

    def alloc_batch_u8(self, n):
        return self.base.alloc_batch_u8(n)


def _shard_file(resume_dir, lo: int, hi: int, n: int, length: int):
    return pathlib.Path(resume_dir) / f'shard_{lo:07d}_{hi:07d}_of_{n}_len{length}.pt'


def predict_sharded(decoder, dataset, world: int = 1, rank: int = 0, batch_size: int = 16, resume_dir=None,
                    **kwargs) -> Sequence[str]:
    """`Decoder.predict` over this rank's shard + all-gather of the token ids; returns all captions.

    With `resume_dir` every rank saves its finished shard (token ids, keyed by the neuron range, so the file is valid
    for any later world size that produces the same range) and a re-run loads it instead of describing the shard
    again: after a rank failure only the missing shards are recomputed (failure detection).
    """
    n = len(dataset)
    if world == 1 and resume_dir is None:
        return decoder.predict(dataset, batch_size=batch_size, **kwargs)
    lo, hi = shard_range(n, rank, world)
    # shards start on reference-batch boundaries only if ceil(N/G) is a multiple of batch_size; the reference's
    # batch-level early exit only changes trailing <stop> columns, never a caption, so captions are unaffected.
    shard = (_SliceU8 if hasattr(dataset, 'batch_u8') else _Slice)(dataset, lo, hi)
    length = kwargs.get('length') or decoder.length
    stop = decoder.indexer.stop_index
    tokens = torch.full((hi - lo, length), stop, dtype=torch.long)
    kwargs.setdefault('display_progress_as', None)
    saved = _shard_file(resume_dir, lo, hi, n, length) if resume_dir is not None else None
    if saved is not None and saved.exists():
        ids = torch.load(saved)
        if tuple(ids.shape) != (hi - lo, length):
            raise RuntimeError(f'rank {rank}: {saved} holds {tuple(ids.shape)}, expected {(hi - lo, length)}')
        tokens = ids
    elif hi > lo:
        captions_local: List[str] = list(decoder.predict(shard, batch_size=batch_size, **kwargs))
        # captions -> token ids would need the tokenizer; gather the ids the engine produced instead
        ids = getattr(decoder, 'last_predict_tokens', None)
        if ids is None or len(ids) != hi - lo or len(captions_local) != hi - lo:
            raise RuntimeError(f'rank {rank}: predict() returned {len(captions_local)} captions and '
                               f'{None if ids is None else len(ids)} token rows for a shard of {hi - lo} neurons')
        tokens[:, :ids.shape[1]] = ids.cpu()
        if saved is not None:  # write-then-rename: a killed rank never leaves a truncated shard behind
            saved.parent.mkdir(parents=True, exist_ok=True)
            partial = saved.with_suffix(f'.tmp{os.getpid()}')
            torch.save(tokens, partial)
            os.replace(partial, saved)
    device = decoder.engine.device if torch.cuda.is_available() else torch.device('cpu')
    gathered = gather_token_rows(tokens.to(device), n, world, stop)
    return tuple(decoder.indexer.reconstruct(gathered.cpu().tolist())) if n else ()
