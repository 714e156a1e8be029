"""How the hot path shards over one-process-per-GPU ranks.

* Eikonal fields (station x phase x model) are independent: the library deals the fields of a slowness
  model over the ranks that hold it (``assign_fields`` = ``mceik_fsm_assign_fields``), heaviest first when
  the iteration counts of an earlier solve are known.  No communication during the solve (this replaces
  the per-sweep ghost exchange of fsm3d.f90:971-1045).
* Every rank writes its fp32 tables straight into its rows of the replicated table buffer and ONE in-place
  NCCL all-gather completes it (``mceik_fsm_solve_sharded_dev``, csrc/comm.cu); the torch helpers below
  are the gloo-testable mirror of that layout.
* Events are independent: contiguous event ranges per rank, tables replicated, no data-path
  collective; the per-event results are gathered once at the end.

``torch.distributed`` supplies the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def block_range(n, world, rank):
    """Contiguous block [lo, hi) of n items for `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_fields(nfields, world, rank):
    """Field indices solved by this rank."""
    lo, hi = block_range(nfields, world, rank)
    return np.arange(lo, hi, dtype=np.int64)


def assign_fields(field_model, world, cost=None):
    """``mceik_fsm_assign_fields``: (rank, table row, rows per rank) of every field -- fields of one slowness model
    dealt over the ranks that hold it, heaviest first when a cost estimate (iteration counts of an earlier solve)
    is given.  The same on every rank."""
    import ctypes as C
    from . import _lib
    fm = np.ascontiguousarray(field_model, dtype=np.int32)
    n = fm.size
    rk, row, slots = np.zeros(n, np.int32), np.zeros(n, np.int32), C.c_int(0)
    c = None if cost is None else np.ascontiguousarray(cost, dtype=np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(_lib.c_int_p)
    _lib.check(_lib.load().mceik_fsm_assign_fields(n, p(fm), p(c), int(world), p(rk), p(row), C.byref(slots)),
               "mceik_fsm_assign_fields")
    return rk, row, slots.value


def shard_events(obs_ptr, world, rank):
    """CSR slice of the events located by this rank -> (event_lo, event_hi, local_obs_ptr, pick_lo, pick_hi)."""
    obs_ptr = np.asarray(obs_ptr)
    lo, hi = block_range(obs_ptr.size - 1, world, rank)
    p0, p1 = int(obs_ptr[lo]), int(obs_ptr[hi])
    return lo, hi, (obs_ptr[lo:hi + 1] - p0).astype(np.int32), p0, p1


def all_gather_tables(local_tables, nfields_total, group=None):
    """Replicate source-sharded fp32 tables [nlocal, ldtab] -> [nfields_total, ldtab] on every rank.

    Ranks may own different numbers of fields (block_range); shorter shards are padded to the
    longest so a single all_gather_into_tensor moves everything (one collective per model)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ld = local_tables.shape[1]
    counts = [block_range(nfields_total, world, r)[1] - block_range(nfields_total, world, r)[0] for r in range(world)]
    assert local_tables.shape[0] == counts[dist.get_rank(group)]
    mx = max(counts)
    if all(c == mx for c in counts):
        out = torch.empty((nfields_total, ld), dtype=local_tables.dtype, device=local_tables.device)
        dist.all_gather_into_tensor(out, local_tables.contiguous(), group=group)
        return out
    padded = torch.zeros((mx, ld), dtype=local_tables.dtype, device=local_tables.device)
    padded[:local_tables.shape[0]] = local_tables
    buf = torch.empty((world * mx, ld), dtype=local_tables.dtype, device=local_tables.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


def gather_event_results(local, nevents_total, group=None):
    """Concatenate per-rank event results (1-D tensors in event order) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [block_range(nevents_total, world, r)[1] - block_range(nevents_total, world, r)[0] for r in range(world)]
    mx = max(counts)
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    buf = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)])
