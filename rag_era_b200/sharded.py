"""Row-sharded multi-GPU plumbing: one process per GPU, ``torch.distributed``
for rendezvous only. The data path is libragera's own: each rank scores its shard and rescoring
its survivors exactly; the ranks' [B][k] exact records are exchanged INSIDE the final fusion kernel
through peer-to-peer mailboxes (CUDA IPC over NVLink), then every rank runs the same merge (rank 0
is the consumer). Bootstrap is host-driven: each rank exports a 64-byte mailbox handle, the handles
are all-gathered here, each rank imports them — no NCCL in the library. ``RAGERA_COMM=nccl`` (or
GPUs without peer access) selects the fallback: one ``ncclAllGather`` of the records before the
fusion kernel, on the library's own communicator.

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


def balanced_ranges(total_rows: int, rows_per_rank: list[int], ms_per_rank: list[float], align: int = 256) -> list[tuple[int, int]]:
    """Contiguous row shards sized by MEASURED speed: rank i scanned rows_per_rank[i] rows in ms_per_rank[i]; the new
    shard sizes are proportional to rows/ms (the batch ends with the slowest shard, so equal TIMES, not equal rows, is
    what minimises it — GPUs of one box differ by several percent under the power cap). Shards stay contiguous, only
    the boundaries (and with them each rank's id_base) move; sizes are multiples of `align` rows except the last."""
    world = len(rows_per_rank)
    speed = [max(r, 1) / max(t, 1e-9) for r, t in zip(rows_per_rank, ms_per_rank)]
    tot = sum(speed)
    out, base = [], 0
    for i in range(world):
        if i == world - 1:
            n = total_rows - base
        else:
            n = int(round(total_rows * speed[i] / tot / align)) * align
            n = max(0, min(n, total_rows - base))
        out.append((base, n))
        base += n
    return out


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


def allgather_bytes(dist, payload: bytes, device=None) -> list[bytes]:
    """Every rank's ``payload`` (equal lengths), ordered by rank — the host channel of the mailbox bootstrap."""
    import torch

    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


def join_exchange(dist, idx: VectorIndex, device: int, max_batch: int = 1024, max_k: int = 64) -> str:
    """Bootstrap the exchange of this rank's index with its peers. Returns "p2p" or "nccl"."""
    import os

    import torch
    from . import _native as N

    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else None
    if os.environ.get("RAGERA_COMM", "p2p") != "nccl":
        handles = allgather_bytes(dist, idx.comm_p2p_export(world, rank, max_batch, max_k), dev)
        try:
            idx.comm_p2p_import(handles)
            ok = 1
        except N.RagError as e:
            if e.code != N.ERR_UNSUPPORTED:
                raise
            ok = 0
        if all(b[0] for b in allgather_bytes(dist, bytes([ok]), dev)):      # every rank mapped every peer
            return "p2p"
        idx.comm_destroy()
    idx.comm_init(world, rank, broadcast_unique_id(dist, rank, dev))
    return "nccl"


def leave_exchange(dist, idx: VectorIndex):
    """Two-phase teardown: every rank unmaps its peers' mailboxes, barrier, then the mailboxes may be freed
    (freeing exported memory a peer still maps is undefined in CUDA)."""
    idx.comm_detach()
    dist.barrier()
    idx.comm_destroy()


def create_sharded_index(dist, total_rows: int, dim: int, dtype: int, device: int, bf16_shadow: bool = False,
                         rows: tuple[int, int] | None = None, max_batch: int = 1024, max_k: int = 64,
                         shadow: str | None = None) -> VectorIndex:
    """Create this rank's shard ([base, base+n) = ``rows`` or the even split) and join the exchange."""
    world, rank = dist.get_world_size(), dist.get_rank()
    base, n = rows if rows is not None else shard_range(total_rows, world, rank)
    idx = VectorIndex(dim, max(n, 1), dtype=dtype, device=device, bf16_shadow=bf16_shadow, id_base=base, shadow=shadow)
    idx.exchange = join_exchange(dist, idx, device, max_batch, max_k) if world > 1 else "none"
    return idx


def open_sharded_cache(dist, cache_path: str, device: int, bf16_shadow: bool = False, shadow: str | None = None):
    """Bring a row-sharded index up from ONE binary sidecar (store_cache.cu): every rank reads the header, creates its
    shard and streams only its own row range (verified block by block) into HBM. Returns (index, all node ids)."""
    from . import _native as N

    info = N.cache_info(cache_path)
    idx = create_sharded_index(dist, info.rows, info.dim, info.dtype, device, bf16_shadow=bf16_shadow, shadow=shadow)
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
