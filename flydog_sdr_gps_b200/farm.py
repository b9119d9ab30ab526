"""Multi-GPU plumbing: one process per GPU, independent units sharded across ranks, only the small
record arrays gathered.  No data-path collective exists on this path: the (capture, sat,
Doppler) tiles share nothing but read-only inputs, and the only cross-tile step -- best-over-Doppler per
(capture, sat) -- stays on the GPU that owns the pair.

  shard by capture (receiver farm, cfg5): every rank holds all code spectra and searches its captures
  shard by satellite (one capture, cfg4) : the capture is given to every rank, each searches a slice of the table

`search_fn` is what runs on the local device (AcqEngine.search in production; the CPU suite substitutes
the oracle to test this plumbing under gloo).
"""
import numpy as np

from .engine import RECORD_DTYPE
from .scenarios import shard


def _gather_records(local, counts, dist, device):
    """all_gather of variable-length record arrays as padded uint8 tensors; returns list per rank."""
    import torch
    world = dist.get_world_size()
    width = max(counts) * RECORD_DTYPE.itemsize
    buf = np.zeros(width, np.uint8)
    raw = local.reshape(-1).view(np.uint8)
    buf[:raw.size] = raw
    t = torch.from_numpy(buf).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.cpu().numpy()[:counts[r] * RECORD_DTYPE.itemsize].view(RECORD_DTYPE) for r, o in enumerate(out)]


def search_sharded_by_capture(search_fn, captures, n_sel, dist=None, device="cpu"):
    """captures: uint8 [n_cap, bytes].  Returns records [n_cap, n_sel] (on every rank)."""
    n_cap = captures.shape[0]
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    lo, hi = shard(n_cap, rank, world)
    local = search_fn(captures[lo:hi]) if hi > lo else np.zeros((0, n_sel), RECORD_DTYPE)
    if world == 1:
        return local
    counts = [(shard(n_cap, r, world)[1] - shard(n_cap, r, world)[0]) * n_sel for r in range(world)]
    parts = _gather_records(local, counts, dist, device)
    return np.concatenate([p.reshape(-1, n_sel) for p in parts], axis=0)


def search_sharded_by_sat(search_fn, capture, n_sats, dist=None, device="cpu"):
    """One capture, satellite table split across ranks.  search_fn(capture, sel) -> records [len(sel)].
    Returns records [n_sats] in table order (on every rank)."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    lo, hi = shard(n_sats, rank, world)
    local = search_fn(capture, np.arange(lo, hi, dtype=np.int32)) if hi > lo else np.zeros(0, RECORD_DTYPE)
    if world == 1:
        return local
    counts = [shard(n_sats, r, world)[1] - shard(n_sats, r, world)[0] for r in range(world)]
    return np.concatenate(_gather_records(local, counts, dist, device))
