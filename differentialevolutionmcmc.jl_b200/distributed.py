"""sample() for a job whose groups are sharded over the GPUs of one box: one process per GPU
(`torchrun --nproc-per-node N script.py`), every rank calls `distributed.sample(model, de, n_iter)`,
rank 0 gets the Chains.  Groups g*G/N .. (g+1)*G/N - 1 live on rank g; migration (src/migration.jl:11-116)
crosses NVLink through NCCL send/recv only when a cycle spans ranks (csrc/kernels.cu: comm_exchange); the
draw map is keyed by global particle positions, so the chains do not depend on N.

torch.distributed is the launch plumbing (rendezvous, broadcast of the NCCL id and of the initial
state, gather of the per-rank histories); the population step itself never touches it."""
from __future__ import annotations

import numpy as np

from .api import DE, DEModel, MCMCThreads, _draw_states, _flatten, build_handle, bundle_samples, resample
from .handle import comm_unique_id


last_counters = None


def _gather_arrays(dist, part, rank, world, device):
    """Every rank's tuple of equally shaped numpy arrays -> on rank 0, a tuple of arrays concatenated along axis 1 (the
    particle positions, in rank order).  With the NCCL backend they travel as device tensors over NVLink and STAY on
    rank 0's GPU as torch tensors (pickling 80 MB per rank through gather_object took longer than the run); the gloo
    backend of the CPU tests gathers the objects."""
    if dist.get_backend() != "nccl":
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0)
        if rank != 0:
            return None
        return tuple(np.concatenate([p[i] for p in parts], axis=1) for i in range(len(part)))
    import torch
    dev = torch.device("cuda", device)
    out = []
    for a in part:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0)
        if rank == 0:
            out.append(torch.cat(bufs, dim=1))
    return tuple(out) if rank == 0 else None


def _merge_by_id(th, w, ids, acc, n_iter, P, d):
    """de.samples[row, :, id] (utilities.jl:161-180): ids travel with the particles through migration.  numpy on the
    host, or -- when the gathered history is still on rank 0's GPU -- torch index_put there and one download."""
    if isinstance(th, np.ndarray):
        rows = np.arange(n_iter)[:, None]
        samples = np.empty((P, d, n_iter))
        lp = np.empty((P, n_iter))
        accept = np.empty((P, n_iter), dtype=np.uint8)
        samples[ids, :, rows] = th
        lp[ids, rows] = w
        accept[ids, rows] = acc
        return samples, lp, accept, ids[-1]
    import torch
    idl = ids.long()
    rows = torch.arange(n_iter, device=th.device)[:, None].expand_as(idl)
    samples = torch.empty((P, d, n_iter), dtype=th.dtype, device=th.device)
    samples[idl, :, rows] = th
    lp = torch.empty((P, n_iter), dtype=w.dtype, device=th.device)
    lp[idl, rows] = w
    accept = torch.empty((P, n_iter), dtype=acc.dtype, device=th.device)
    accept[idl, rows] = acc
    return samples.cpu().numpy(), lp.cpu().numpy(), accept.cpu().numpy(), ids[-1].cpu().numpy()


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
            box = [_draw_states(model.sample_prior, P, d) if rank == 0 else None]
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
    th, w, ids, acc = parts                                  # [n][P][d], [n][P] x 3 -- by position, ranks side by side
    if not isinstance(th, np.ndarray):
        return _bundle_on_gpu(model, de, th, w, ids, acc, init_rows, shapes, n_iter, P, d)
    samples, lp, accept, final_ids = _merge_by_id(th, w, ids, acc, n_iter, P, d)
    if init_rows is not None:
        # rows 1..n_initial of de.samples are the prior draws (utilities.jl:35-39); bundle_samples then keeps
        # rows burnin+1..n_iter of the n_iter + n_initial array -- not shifted by n_initial (main.jl:226-234)
        n0 = de.n_initial
        samples = np.concatenate([init_rows.transpose(1, 2, 0), samples], axis=2)
        lp = np.concatenate([np.zeros((P, n0)), lp], axis=1)
        accept = np.concatenate([np.zeros((P, n0), dtype=np.uint8), accept], axis=1)
    de.samples = samples
    return bundle_samples(model, de, samples, accept, lp, final_ids, shapes, n_iter)


_stage = [None]                                              # page-locked staging tensor, kept for the life of the process


def _download(t):
    """GPU tensor -> fresh numpy array through a cached page-locked staging block (torch's .cpu() into pageable memory
    ran at 2 GB/s: 8 ms for the 17 MB bundle of a 20-iteration, 2-GPU run); 16 MB chunks, the copy of chunk c into the
    array overlapping the transfer of chunk c + 1."""
    import torch
    flat = t.contiguous().view(-1)
    out = np.empty(tuple(t.shape), dtype=np.float64)
    n = flat.numel()
    if n == 0:
        return out
    chunk = 2 << 20                                         # doubles per staging half: 16 MB
    if _stage[0] is None:
        _stage[0] = torch.empty(2 * chunk, dtype=torch.float64, pin_memory=True)
    st, dst = _stage[0], out.reshape(-1)
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    spans = [(o, min(chunk, n - o)) for o in range(0, n, chunk)]
    for c, (o, m) in enumerate(spans[:1]):
        st[:m].copy_(flat[o:o + m], non_blocking=True); ev[0].record()
    for c, (o, m) in enumerate(spans):
        if c + 1 < len(spans):
            o2, m2 = spans[c + 1]
            h2 = ((c + 1) & 1) * chunk
            st[h2:h2 + m2].copy_(flat[o2:o2 + m2], non_blocking=True); ev[(c + 1) & 1].record()
        ev[c & 1].synchronize()
        h = (c & 1) * chunk
        dst[o:o + m] = st[h:h + m].numpy()
    return out


def _bundle_on_gpu(model, de, th, w, ids, acc, init_rows, shapes, n_iter, P, d):
    """The by-id merge and bundle_samples (src/main.jl:222-250, by-position quirk of "acceptance" / "lp" included) as
    index operations on rank 0's GPU, where the gathered history already is; ONE download of the array the Chains wrap,
    in the memory order the single-GPU sample() returns (chain, column, draw) so no host-side transpose is needed."""
    import torch
    from .api import Chains, _flat_names
    dev = th.device
    idl = ids.long()
    rows = torch.arange(n_iter, device=dev)[:, None].expand_as(idl)
    n0 = de.n_initial if init_rows is not None else 0
    samples = torch.empty((P, d, n0 + n_iter), dtype=th.dtype, device=dev)
    lp = torch.zeros((P, n0 + n_iter), dtype=w.dtype, device=dev)
    accept = torch.zeros((P, n0 + n_iter), dtype=torch.float64, device=dev)
    samples[idl, :, rows + n0] = th
    lp[idl, rows + n0] = w
    accept[idl, rows + n0] = acc.double()
    if n0:
        samples[:, :, :n0] = torch.from_numpy(np.ascontiguousarray(init_rows.transpose(1, 2, 0))).to(dev)
    Ns = n_iter - de.burnin if de.discard_burnin else n_iter
    offset = de.burnin if de.discard_burnin else 0
    Ns = max(Ns, 0)
    arr = torch.empty((P, d + 2, Ns), dtype=th.dtype, device=dev)
    if Ns > 0:
        final_ids = idl[-1]
        arr[:, :d, :] = samples[:, :, offset:offset + Ns]
        arr[:, d, :] = accept[final_ids, offset:offset + Ns]
        arr[:, d + 1, :] = lp[final_ids, offset:offset + Ns]
    host = _download(arr)
    de.samples = host[:, :d, :]
    names = _flat_names(model.names, shapes) + ["acceptance", "lp"]
    return Chains(host.transpose(2, 1, 0), names, [str(n) for n in model.names])
