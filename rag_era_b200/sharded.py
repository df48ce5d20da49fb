"""Row-sharded multi-GPU plumbing: one process per GPU, ``torch.distributed``
for rendezvous only. The data path is libragera's own: each rank scores its shard, rescoring
its survivors exactly, one ``ncclAllGather`` of the [B][k] exact records on the library
stream, then every rank runs the same K5 merge (rank 0 is the consumer).

The reference has no counterpart (single Node process); this is the build's addition.
"""
from __future__ import annotations

import numpy as np

from .index import VectorIndex


def shard_range(total_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous row shard of rank ``rank``: rows [base, base+n). Chunk id = base + local row, so
    "lower id wins ties" is independent of the shard count."""
    per = (total_rows + world_size - 1) // world_size
    base = min(total_rows, rank * per)
    return base, max(0, min(per, total_rows - base))


def broadcast_unique_id(dist, rank: int, device=None) -> bytes:
    """Rank 0 creates the NCCL unique id; ``dist.broadcast`` hands it to the other ranks."""
    import torch

    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(VectorIndex.comm_unique_id()), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def create_sharded_index(dist, total_rows: int, dim: int, dtype: int, device: int, bf16_shadow: bool = False) -> VectorIndex:
    """Create this rank's shard and join the library communicator."""
    world, rank = dist.get_world_size(), dist.get_rank()
    base, n = shard_range(total_rows, world, rank)
    idx = VectorIndex(dim, max(n, 1), dtype=dtype, device=device, bf16_shadow=bf16_shadow, id_base=base)
    if world > 1:
        import torch

        uid = broadcast_unique_id(dist, rank, torch.device("cuda", device) if dist.get_backend() == "nccl" else None)
        idx.comm_init(world, rank, uid)
    return idx


def open_sharded_cache(dist, cache_path: str, device: int, bf16_shadow: bool = False):
    """Bring a row-sharded index up from ONE binary sidecar (store_cache.cu): every rank reads the header, creates its
    shard and streams only its own row range (verified block by block) into HBM. Returns (index, all node ids)."""
    from . import _native as N

    info = N.cache_info(cache_path)
    idx = create_sharded_index(dist, info.rows, info.dim, info.dtype, device, bf16_shadow=bf16_shadow)
    base, n = shard_range(info.rows, dist.get_world_size(), dist.get_rank())
    ids = idx.load_cache(cache_path, first_row=base, nrows=n) if n else []
    return idx, ids


def merge_reference_order(ids_per_shard, scores_per_shard, k: int):
    """The order every rank's K5 merge implements, stated on host arrays for tests of the
    exchange protocol: (score desc, chunk id asc), first k. Not used by the product path."""
    ids = np.concatenate([np.asarray(a, dtype=np.uint64) for a in ids_per_shard])
    sc = np.concatenate([np.asarray(a, dtype=np.float64) for a in scores_per_shard])
    order = np.lexsort((ids, -sc))[:k]
    return ids[order], sc[order]
