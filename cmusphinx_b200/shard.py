"""Host-side multi-GPU plumbing: one process per GPU, utterance / frame shards,
ONE broadcast of the packed model parameters at load, no collective on the
per-frame path.

The sharding mirrors pocketsphinx_batch's own process-level sharding options
(-ctloffset / -ctlcount / -ctlincr, pocketsphinx/src/programs/batch.c:73-81,
702-708): rank r of n takes control-file lines r, r+n, r+2n, ... (ctlincr = n,
ctloffset = r) or a contiguous block.
"""
import hashlib
from typing import Dict, List, Sequence, Tuple

import numpy as np

PARAM_ORDER = ("mean", "var", "det", "mixw")


def shard_strided(n_items: int, rank: int, world: int) -> range:
    """-ctloffset rank -ctlincr world."""
    return range(rank, n_items, world)


def shard_block(n_items: int, rank: int, world: int) -> range:
    """Contiguous -ctloffset/-ctlcount block; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def pack_params(params: Dict[str, np.ndarray]) -> Tuple[np.ndarray, List[Tuple[str, str, Tuple[int, ...]]]]:
    """Flattens the precomputed model arrays into one uint8 blob + a manifest."""
    manifest, chunks = [], []
    for k in PARAM_ORDER:
        a = np.ascontiguousarray(params[k])
        manifest.append((k, a.dtype.str, a.shape))
        chunks.append(a.view(np.uint8).reshape(-1))
    return np.concatenate(chunks), manifest


def unpack_params(blob: np.ndarray, manifest) -> Dict[str, np.ndarray]:
    out, off = {}, 0
    for k, dt, shape in manifest:
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        out[k] = blob[off:off + n].view(np.dtype(dt)).reshape(shape).copy()
        off += n
    return out


def checksum(blob: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(blob).tobytes()).hexdigest()


def broadcast_params(params, src: int = 0, device=None):
    """Rank `src` passes the dict of arrays, the others None; everyone returns
    the same dict.  One broadcast of the manifest (object) and one of the blob
    (tensor; NCCL when `device` is a CUDA device, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    if rank == src:
        blob, manifest = pack_params(params)
        meta = [manifest, int(blob.size)]
    else:
        blob, meta = None, [None, None]
    if device is not None:
        dist.broadcast_object_list(meta, src=src, device=device)
    else:
        dist.broadcast_object_list(meta, src=src)
    manifest, n = meta
    t = torch.from_numpy(blob) if rank == src else torch.empty(n, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    blob = t.cpu().numpy()
    return unpack_params(blob, manifest), checksum(blob)


def gather_hypotheses(local: Sequence[Tuple[int, str]], dst: int = 0):
    """Host-side gather of (utterance index, hypothesis) pairs; rank dst gets
    them merged in utterance order, the others None."""
    import torch.distributed as dist
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(list(local), bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = [p for part in bucket for p in part]
    return [h for _, h in sorted(merged)]
