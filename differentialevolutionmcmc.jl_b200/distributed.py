"""sample() for a job whose groups are sharded over the GPUs of one box: one process per GPU
(`torchrun --nproc-per-node N script.py`), every rank calls `distributed.sample(model, de, n_iter)`,
rank 0 gets the Chains.  Groups g*G/N .. (g+1)*G/N - 1 live on rank g; migration (src/migration.jl:11-116)
crosses NVLink through NCCL send/recv only when a cycle spans ranks (csrc/kernels.cu: comm_exchange); the
draw map is keyed by global particle positions, so the chains do not depend on N.

torch.distributed is the launch plumbing (rendezvous, broadcast of the NCCL id and of the initial
state, gather of the per-rank histories); the population step itself never touches it."""
from __future__ import annotations

import numpy as np

from .api import DE, DEModel, MCMCThreads, _flatten, build_handle, bundle_samples, resample
from .handle import comm_unique_id


last_counters = None


def _gather_arrays(dist, part, rank, world, device):
    """Every rank's tuple of equally shaped numpy arrays -> list of tuples on rank 0.  With the NCCL backend the arrays
    travel as device tensors over NVLink (pickling 80 MB per rank through gather_object took longer than the run);
    the gloo backend of the CPU tests gathers the objects."""
    if dist.get_backend() != "nccl":
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0)
        return parts
    import torch
    dev = torch.device("cuda", device)
    out = [[] for _ in range(world)]
    for a in part:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0)
        if rank == 0:
            for r in range(world):
                out[r].append(bufs[r].cpu().numpy())
    return [tuple(o) for o in out] if rank == 0 else None


def sample(model: DEModel, de: DE, *args, device=None, unique_id=None):
    """sample(model, de, n_iter) / sample(model, de, MCMCThreads(), n_iter) on all ranks of the default
    process group.  Returns the Chains on rank 0 and None elsewhere.  `unique_id`: a communicator id
    to use instead of a fresh NCCL one (the tests' host-only double)."""
    import torch.distributed as dist

    if len(args) == 2 and isinstance(args[0], MCMCThreads):
        n_iter = int(args[1])
    elif len(args) == 1:
        n_iter = int(args[0])
    else:
        raise TypeError("sample(model, de, n_iter) or sample(model, de, MCMCThreads(), n_iter)")
    if not dist.is_initialized():
        raise RuntimeError("distributed.sample needs an initialised torch.distributed process group (torchrun)")
    rank, world = dist.get_rank(), dist.get_world_size()
    if de.n_groups % world:
        raise ValueError(f"n_groups = {de.n_groups} does not divide over {world} ranks")
    init_rows = None
    per = de.n_groups // world
    P, P_local = de.n_groups * de.Np, per * de.Np
    if de.seed is None:                                     # every rank must plan the same migrations
        box = [int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0]) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        de.seed = box[0]
    h, shapes, d = build_handle(model, de, device=rank if device is None else device, group_begin=rank * per, group_count=per, n_iter=n_iter)
    try:
        box = [(unique_id if unique_id is not None else comm_unique_id()) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        h.comm_init(box[0], rank, world)
        if de.n_initial > 0:
            # initialize_samples (src/utilities.jl:29-41): n_initial sample_prior() draws per particle id, drawn
            # once on rank 0; every rank keeps ALL of them (resample draws donors from every id's history) and
            # init_particle starts each particle from samples[1, :, id] (utilities.jl:15)
            rows = None
            if rank == 0:
                rows = np.empty((de.n_initial, P, d))
                for p in range(P):
                    for i in range(de.n_initial):
                        rows[i, p] = _flatten(model.sample_prior())
            box = [rows]
            dist.broadcast_object_list(box, src=0)
            init_rows = box[0]
            if de.sample is resample:
                h.set_history(box[0])
            else:
                h.set_history(box[0][:, rank * P_local:(rank + 1) * P_local])
            h.set_state(None)
        else:
            # sample_init (src/main.jl:263-271): one sample_prior() per particle in id order -- drawn once, on rank 0
            box = [np.array([_flatten(model.sample_prior()) for _ in range(P)], dtype=np.float64) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            h.set_state(box[0][rank * P_local:(rank + 1) * P_local])
        h.run(n_iter)
        de.iter = n_iter + de.n_initial
        global last_counters
        last_counters = h.counters()                         # of this rank's shard (which migration transport ran, launches, ...)
        part = h.history_by_slot()                           # theta[n][P_local][d], w, ids, acc -- by position
        parts = _gather_arrays(dist, part, rank, world, h.cfg.device)
    finally:
        h.close()
    if rank != 0:
        return None
    th = np.concatenate([p[0] for p in parts], axis=1)       # [n][P][d]
    w = np.concatenate([p[1] for p in parts], axis=1)
    ids = np.concatenate([p[2] for p in parts], axis=1)
    acc = np.concatenate([p[3] for p in parts], axis=1)
    # de.samples[row, :, id] (utilities.jl:161-180): ids travel with the particles through migration
    rows = np.arange(n_iter)[:, None]
    samples = np.empty((P, d, n_iter))
    lp = np.empty((P, n_iter))
    accept = np.empty((P, n_iter), dtype=np.uint8)
    samples[ids, :, rows] = th
    lp[ids, rows] = w
    accept[ids, rows] = acc
    if init_rows is not None:
        # rows 1..n_initial of de.samples are the prior draws (utilities.jl:35-39); bundle_samples then keeps
        # rows burnin+1..n_iter of the n_iter + n_initial array -- not shifted by n_initial (main.jl:226-234)
        n0 = de.n_initial
        samples = np.concatenate([init_rows.transpose(1, 2, 0), samples], axis=2)
        lp = np.concatenate([np.zeros((P, n0)), lp], axis=1)
        accept = np.concatenate([np.zeros((P, n0), dtype=np.uint8), accept], axis=1)
    de.samples = samples
    return bundle_samples(model, de, samples, accept, lp, ids[-1], shapes, n_iter)
