"""N > 1 path on CPU: two processes (torch.distributed, gloo), each holding half of the groups
through the host-only test double, exchange migrating particles; the gathered result must equal
the single-process run bit for bit (the draw map is keyed by global positions, migration moves
theta/weight/id/accept between ranks).  On the GPU the same exchange is NCCL send/recv
(kernels.cu: comm_exchange)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import common

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["DEMCMC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DEMCMC_ROOT"], "tests"))
import common
from common import D, make_case
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
common.use_emu()
L = D._ffi.lib()
FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p)
def exchange(rk, n, src, dst, send, recv, row_len, user):
    reqs, bufs = [], []
    for i in range(n):
        if src[i] == dst[i]:
            continue
        r = (i + n - 1) % n
        if rk == src[i]:
            t = torch.from_numpy(np.ctypeslib.as_array(send, shape=(n * row_len,))[r * row_len:(r + 1) * row_len].copy())
            reqs.append(dist.isend(t, dst[i], tag=i))
        if rk == dst[i]:
            t = torch.empty(row_len, dtype=torch.float64)
            bufs.append((r, t))
            reqs.append(dist.irecv(t, src[i], tag=i))
    for q in reqs:
        q.wait()
    out = np.ctypeslib.as_array(recv, shape=(n * row_len,))
    for r, t in bufs:
        out[r * row_len:(r + 1) * row_len] = t.numpy()
    return 0
cb = FN(exchange)
L.demcmc_emu_set_exchange(cb, None)
G, Np, n_iter = 4, 6, 60
case = make_case("gaussian", np.random.default_rng(31))
theta0 = case.theta0(np.random.default_rng(5), G * Np)
per = G // world
h = case.handle(G, Np, seed=17, burnin=20, alpha=0.5, theta_snooker=0.2, group_begin=rank * per, group_count=per)
h.comm_init(b"\0" * 128, rank, world)
h.set_state(theta0[rank * per * Np:(rank + 1) * per * Np])
h.run(n_iter)
th, w, ids, acc = h.history_by_slot()
mig = h.migration_slots()
h.close()
np.savez(os.environ["DEMCMC_OUT"] + f".{rank}.npz", th=th, w=w, ids=ids, acc=acc, mig=mig)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_equal_one(tmp_path, emu):
    out = str(tmp_path / "mr")
    env = dict(os.environ, DEMCMC_ROOT=common.ROOT, DEMCMC_OUT=out, MASTER_PORT="29571", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, "-c", WORKER], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    parts = [np.load(out + f".{r}.npz") for r in range(2)]
    G, Np, n_iter = 4, 6, 60
    case = common.make_case("gaussian", np.random.default_rng(31))
    theta0 = case.theta0(np.random.default_rng(5), G * Np)
    h = case.handle(G, Np, seed=17, burnin=20, alpha=0.5, theta_snooker=0.2)
    h.set_state(theta0)
    h.run(n_iter)
    th, w, ids, acc = h.history_by_slot()
    mig = h.migration_slots()
    h.close()
    assert (mig >= 0).sum() > 20                                   # migrations happened ...
    assert len({tuple(r) for r in ids}) > 5                        # ... and moved ids around
    assert any(set(parts[0]["ids"][-1]) - set(range(12)))          # ... across the rank boundary
    for k, full in (("th", th), ("w", w), ("ids", ids), ("acc", acc)):
        got = np.concatenate([parts[0][k], parts[1][k]], axis=1)
        assert np.array_equal(got, full), k
    # each rank logged the picks of its own groups
    merged = np.maximum(parts[0]["mig"], parts[1]["mig"])
    assert np.array_equal(merged, mig)


WORKER_API = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["DEMCMC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DEMCMC_ROOT"], "tests"))
import common
from common import D
from demcmc_b200 import distributed
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
common.use_emu()
L = D._ffi.lib()
FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p)
def exchange(rk, n, src, dst, send, recv, row_len, user):
    reqs, bufs = [], []
    for i in range(n):
        if src[i] == dst[i]:
            continue
        r = (i + n - 1) % n
        if rk == src[i]:
            t = torch.from_numpy(np.ctypeslib.as_array(send, shape=(n * row_len,))[r * row_len:(r + 1) * row_len].copy())
            reqs.append(dist.isend(t, dst[i], tag=i))
        if rk == dst[i]:
            t = torch.empty(row_len, dtype=torch.float64)
            bufs.append((r, t))
            reqs.append(dist.irecv(t, src[i], tag=i))
    for q in reqs:
        q.wait()
    out = np.ctypeslib.as_array(recv, shape=(n * row_len,))
    for r, t in bufs:
        out[r * row_len:(r + 1) * row_len] = t.numpy()
    return 0
AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
def allgather(send, recv, nbytes, user):
    mine = torch.from_numpy(np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy())
    parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(nbytes * world,))
    for r, t in enumerate(parts):
        out[r * nbytes:(r + 1) * nbytes] = t.numpy()
    return 0
cb, cb2 = FN(exchange), AG(allgather)
L.demcmc_emu_set_exchange(cb, None)
L.demcmc_emu_set_allgather(cb2, None)
exec(os.environ["DEMCMC_MODEL"])
chains = distributed.sample(model, de, 50, device=0, unique_id=b"\\0" * 128)
if rank == 0:
    np.save(os.environ["DEMCMC_OUT"] + ".npy", chains.value)
dist.barrier()
dist.destroy_process_group()
'''

MODEL = r'''
rng = np.random.default_rng(3)
x = np.random.default_rng(11).normal(0.3, 1.2, 40)
model = D.DEModel(sample_prior=lambda: [rng.normal(), abs(rng.standard_cauchy()) + 0.1], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                  loglike=D.GPULoglike("gaussian", x), names=("mu", "sigma"))
de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=4, Np=5, burnin=20, seed=99, **{"α": 0.5, "θsnooker": 0.2})
'''


MODEL_RESAMPLE = MODEL.replace('n_groups=4, Np=5, burnin=20, seed=99,', 'n_groups=4, Np=5, burnin=20, seed=99, n_initial=6, sample=D.resample,')


@pytest.mark.parametrize("which", ["current", "resample"])
def test_distributed_sample_equals_single_process(tmp_path, emu, which):
    """demcmc_b200.distributed.sample on two ranks (the user-facing call of a sharded job) returns the
    very Chains of the single-process sample(): same sample_prior() draws, same seed, ids followed
    through migrations across the rank boundary, bundle_samples on the gathered history."""
    out = str(tmp_path / "api")
    MODEL_ = MODEL if which == "current" else MODEL_RESAMPLE
    env = dict(os.environ, DEMCMC_ROOT=common.ROOT, DEMCMC_OUT=out, MASTER_PORT="29573" if which == "current" else "29577", WORLD_SIZE="2", OMP_NUM_THREADS="1", DEMCMC_MODEL=MODEL_)
    procs = [subprocess.Popen([sys.executable, "-c", WORKER_API], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    got = np.load(out + ".npy")
    D = common.D
    ns = {"np": np, "D": D}
    exec(MODEL_, ns)
    ref = D.sample(ns["model"], ns["de"], 50)
    assert got.shape == ref.value.shape == (30, 4, 20)
    assert np.array_equal(got, ref.value)


WORKER_RESAMPLE = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["DEMCMC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DEMCMC_ROOT"], "tests"))
import common
from common import D, make_case
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
common.use_emu()
L = D._ffi.lib()
FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p)
def exchange(rk, n, src, dst, send, recv, row_len, user):
    reqs, bufs = [], []
    for i in range(n):
        if src[i] == dst[i]:
            continue
        r = (i + n - 1) % n
        if rk == src[i]:
            t = torch.from_numpy(np.ctypeslib.as_array(send, shape=(n * row_len,))[r * row_len:(r + 1) * row_len].copy())
            reqs.append(dist.isend(t, dst[i], tag=i))
        if rk == dst[i]:
            t = torch.empty(row_len, dtype=torch.float64)
            bufs.append((r, t))
            reqs.append(dist.irecv(t, src[i], tag=i))
    for q in reqs:
        q.wait()
    out = np.ctypeslib.as_array(recv, shape=(n * row_len,))
    for r, t in bufs:
        out[r * row_len:(r + 1) * row_len] = t.numpy()
    return 0
AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
def allgather(send, recv, nbytes, user):
    mine = torch.from_numpy(np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy())
    parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(nbytes * world,))
    for r, t in enumerate(parts):
        out[r * nbytes:(r + 1) * nbytes] = t.numpy()
    return 0
cb, cb2 = FN(exchange), AG(allgather)
L.demcmc_emu_set_exchange(cb, None)
L.demcmc_emu_set_allgather(cb2, None)
G, Np, n_iter, n0 = 4, 6, 40, 7
case = make_case("gaussian", np.random.default_rng(31))
rng = np.random.default_rng(5)
rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)])
per = G // world
h = case.handle(G, Np, seed=17, burnin=15, alpha=0.5, theta_snooker=0.2, n_initial=n0, resample=True, group_begin=rank * per, group_count=per)
h.comm_init(b"\\0" * 128, rank, world)
h.set_history(rows)
h.set_state(None)
h.run(n_iter)
th, w, ids, acc = h.history_by_slot(0, n_iter)
h.close()
np.savez(os.environ["DEMCMC_OUT"] + f".{rank}.npz", th=th, w=w, ids=ids, acc=acc)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_resample_equals_one(tmp_path, emu):
    """sample = resample (DE-MCz, crossover.jl:113-124) on a sharded job: donors are cells of the history of
    ALL particle ids, so every rank keeps a replicated copy of each row (all-gather after each iteration and
    after each migration).  Two ranks must reproduce the single-process chain bit for bit."""
    out = str(tmp_path / "rs")
    env = dict(os.environ, DEMCMC_ROOT=common.ROOT, DEMCMC_OUT=out, MASTER_PORT="29575", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, "-c", WORKER_RESAMPLE], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    parts = [np.load(out + f".{r}.npz") for r in range(2)]
    G, Np, n_iter, n0 = 4, 6, 40, 7
    case = common.make_case("gaussian", np.random.default_rng(31))
    rng = np.random.default_rng(5)
    rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)])
    h = case.handle(G, Np, seed=17, burnin=15, alpha=0.5, theta_snooker=0.2, n_initial=n0, resample=True)
    h.set_history(rows)
    h.set_state(None)
    h.run(n_iter)
    th, w, ids, acc = h.history_by_slot(0, n_iter)
    h.close()
    assert any(set(parts[0]["ids"][-1]) - set(range(12)))          # particles crossed the rank boundary
    for k, full in (("th", th), ("w", w), ("ids", ids), ("acc", acc)):
        got = np.concatenate([parts[0][k], parts[1][k]], axis=1)
        assert np.array_equal(got, full), k
