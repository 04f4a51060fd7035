"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).

The render-and-compare path shards by hypothesis: rows of the (detection, hypothesis) table are independent until the
top-K, objects/instances are independent throughout.  So:

  * coarse stage   every rank scores a contiguous slice of the B*M rows; ONE all-gather of the fp32 logits
                   ([B*M] floats, e.g. 240 x 576 x 4 B = 553 KB) gives every rank the full score table; the
                   segmented top-K is then replicated (deterministic, so no index exchange).
  * refiner stage  the surviving rows are sliced the same way; each row's iterations stay on one GPU; one all-gather
                   of (pose 16 x f32) and one of the scoring logits at the end.

The reference has no such collective: it shards whole scenes across ranks and gathers predictions through pickle
files in a tmp dir plus barriers (toolbox/utils/tensor_collection.py:166-187, toolbox/utils/distributed.py:46-77).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def init_distributed_mode(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun); binds this process to cuda:LOCAL_RANK.
    Unlike the reference (toolbox/utils/distributed.py:131-152) no CUDA_VISIBLE_DEVICES rewriting is needed."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shard_bounds(n: int, world: int) -> List[int]:
    """Contiguous slices of size ceil(n/world) (the last ones may be short or empty): keeps one detection's
    hypotheses on at most two GPUs."""
    per = (n + world - 1) // world if world > 0 else n
    return [min(n, r * per) for r in range(world + 1)]


def shard_range(n: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    b = shard_bounds(n, world)
    return b[rank], b[rank + 1]


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Inverse of shard_range: every rank passes its slice [hi-lo, ...] and gets the full [n_total, ...] tensor.
    One all_gather on equal-size (padded) buffers -- the collective of the coarse stage."""
    world = get_world_size()
    if world == 1:
        assert local.shape[0] == n_total
        return local
    bounds = shard_bounds(n_total, world)
    per = bounds[1] - bounds[0] if world > 0 else n_total
    per = max(per, 1)
    buf = local.new_zeros((per,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    gathered = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(gathered, buf.contiguous())
    parts = [gathered[r * per : r * per + (bounds[r + 1] - bounds[r])] for r in range(world)]
    return torch.cat(parts, dim=0)


def all_gather_rows_packed(local: dict, n_total: int) -> dict:
    """all_gather_rows for a whole dict of row-aligned float tensors with ONE collective: the tensors of this rank's rows
    are flattened per row and packed side by side into a [rows, sum(widths)] buffer, gathered, and split again.
    (The refiner stage returns 6 tensors for each of its iterations; one latency-bound collective instead of 30.)"""
    keys = list(local.keys())
    if get_world_size() == 1 or not keys:
        return dict(local)
    rows = local[keys[0]].shape[0]
    tails = [tuple(local[k].shape[1:]) for k in keys]
    widths = [int(torch.Size(t).numel()) for t in tails]
    packed = torch.cat([local[k].reshape(rows, w).float() for k, w in zip(keys, widths)], dim=1)
    full = all_gather_rows(packed.contiguous(), n_total)
    out, off = {}, 0
    for k, w, t in zip(keys, widths, tails):
        out[k] = full[:, off:off + w].reshape((n_total,) + t).to(local[k].dtype).contiguous()
        off += w
    return out


def all_gather_collections(coll):
    """PandasTensorCollection.gather_distributed without tmp files: all_gather_object of the (small) per-rank
    collections; every rank returns the concatenation in rank order."""
    from .utils.tensor_collection import concatenate

    if not is_distributed():
        return coll
    device = coll.device if len(coll.tensors) else torch.device("cpu")
    parts: List[object] = [None] * get_world_size()
    dist.all_gather_object(parts, coll.clone().cpu() if len(coll.tensors) else coll)  # clone: .cpu() moves in place
    out = concatenate(parts)
    return out.to(device) if len(out.tensors) else out


def all_ranks_equal(values, device=None) -> bool:
    """True when the list of floats `values` is identical on every rank (one MIN and one MAX all-reduce, one host read)."""
    if not is_distributed():
        return True
    backend = dist.get_backend()
    dev = torch.device(device) if (backend == "nccl" and device is not None) else torch.device("cpu")
    v = torch.tensor([float(x) for x in values], dtype=torch.float64, device=dev)
    both = torch.stack([v, -v])
    dist.all_reduce(both, op=dist.ReduceOp.MAX)  # max(v) and -min(v)
    return bool(torch.equal(both[0], -both[1]))


def barrier() -> None:
    if is_distributed():
        dist.barrier()
