"""Multi-GPU plumbing for batches of independent pairs.

The path shards trivially: pairs have no cross-dependencies, so each rank (one
process per GPU) owns a contiguous pair-index range.  Exactly two collectives:
one broadcast of the geometry arena per scene and one all-gather of the
fixed-size result records per batch -- both through torch.distributed (NCCL over
NVLink on GPUs; gloo in the CPU tests).  No collective touches the data path of a
pair.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Equal contiguous ranges of ceil(n/world) pairs; the last ranks may be short or empty."""
    per = (n + world - 1) // world if world > 0 else n
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)], per


def _as_bytes_tensor(a, device):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy())
    return t.to(device) if device is not None else t


def broadcast_array(a, dtype, src=0, device=None):
    """Broadcast a numpy (structured) array from `src`; other ranks pass None."""
    rank = dist.get_rank()
    hdr = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0] = np.ascontiguousarray(a).nbytes
    dist.broadcast(hdr, src)
    nbytes = int(hdr.item())
    if rank == src:
        buf = _as_bytes_tensor(a, device)
    else:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf.cpu().numpy().view(dtype)


def broadcast_geometry(shapes, convex_points, src=0, device=None):
    """One broadcast of the scene: shape records + the list of convex vertex sets."""
    from . import _pod as P
    rank = dist.get_rank()
    shapes = broadcast_array(shapes if rank == src else None, P.shape_dtype, src, device)
    if rank == src:
        counts = np.array([len(p) for p in convex_points], dtype=np.int64)
        flat = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in convex_points]) \
            if len(convex_points) else np.zeros((0, 3))
    else:
        counts = flat = None
    counts = broadcast_array(counts, np.int64, src, device)
    flat = broadcast_array(flat, np.float64, src, device).reshape(-1, 3)
    offs = np.concatenate([[0], np.cumsum(counts)])
    return shapes, [flat[offs[i]:offs[i + 1]] for i in range(len(counts))]


def all_gather_records(local, per, n_total, dtype, device=None):
    """All-gather equal-size per-rank record blocks (padded to `per`) and trim to n_total."""
    world = dist.get_world_size()
    item = np.dtype(dtype).itemsize
    pad = np.zeros(per, dtype=dtype)
    pad[:len(local)] = local
    mine = _as_bytes_tensor(pad, device)
    out = torch.empty(per * item * world, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().view(dtype)[:n_total] if n_total <= per * world else None


def sharded_batch(compute, n, dtype, device=None):
    """Run `compute(lo, hi) -> records[lo:hi]` on this rank's range and all-gather the result so
    that every rank ends up with all n records in pair order."""
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds, per = shard_bounds(n, world)
    lo, hi = bounds[rank]
    local = compute(lo, hi) if hi > lo else np.zeros(0, dtype=dtype)
    full = all_gather_records(local, per, per * world, dtype, device)
    # ranks hold [r*per, r*per + len_r): with contiguous equal ranges this is already pair order
    return full[:n]


# ---- GPU path: the product's own communicator (hfb_comm_*, NCCL behind the C-ABI) --------------------------
# torch.distributed is only the launcher here (rank numbers and the exchange of the 128-byte communicator id);
# geometry broadcast and result all-gather are the library's, and nothing below touches host memory.

def init_engine_comm(eng):
    """hfb_comm_init of `eng` over the torch.distributed world: rank 0 creates the id, everybody receives it"""
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(ids[0], rank, world)


def sharded_distance_device(eng, n_local, d_h1, d_tf1, d_h2, d_tf2, req=None, stream=0):
    """this rank's device rows -> device pointer of all ranks' records (rank-major), all-gathered by the library on
    its own stream (overlapping the caller's next batch); eng.comm_wait(stream) before reading it"""
    return eng.batch_distance_sharded_device(n_local, d_h1, d_tf1, d_h2, d_tf2, req, stream)
