"""torchrun --nproc-per-node N scripts/dist_sample_demo.py : the user-facing sharded call,
demcmc_b200.distributed.sample, on N GPUs; rank 0 compares the Chains with the single-GPU sample()
of the same model, seed and sample_prior() draws (bit for bit) and prints the throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import demcmc_b200 as D
from demcmc_b200 import distributed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D._ffi.use_library(D._ffi.DEFAULT_LIB)
n, dm, G, Np, n_iter = 20000, 50, 4 * world, 64, 120
x = np.random.default_rng(5).normal(np.random.default_rng(6).normal(size=dm), 1.0, size=(n, dm))

RESAMPLE = len(sys.argv) > 1 and sys.argv[1] == "resample"


def make():
    rng = np.random.default_rng(3)
    model = D.DEModel(sample_prior=lambda: [rng.normal(size=dm), abs(rng.standard_cauchy()) + 0.2], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                      loglike=D.GPULoglike("mvnormal", x), names=("mu", "sigma"))
    extra = dict(n_initial=12, sample=D.resample) if RESAMPLE else {}     # DE-MCz donors: replicated history, ncclAllGather per iteration
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=G, Np=Np, burnin=40, seed=2026, **{"α": 0.3, "θsnooker": 0.1}, **extra)
    return model, de

model, de = make()
distributed.sample(model, de, 8, device=local)                      # warm-up (contexts, NCCL connections)
model, de = make()
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
chains = distributed.sample(model, de, n_iter, device=local)
dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    model, de = make()
    ref = D.sample(model, de, n_iter, device=local)
    same = np.array_equal(chains.value, ref.value)
    print(f"distributed.sample{' (sample = resample)' if RESAMPLE else ''} on {world} GPUs: chains {chains.value.shape}, identical to the single-GPU sample(): {same}; "
          f"{G * Np * n_iter / dt:.0f} particle-updates/s end to end", flush=True)
    assert same
    c = distributed.last_counters
    print(f"migration transport: DEMCMC_MIG={os.environ.get('DEMCMC_MIG', 'mailbox (default)')}: {c['cross_migrations']} cross-rank migrations, "
          f"{c['mailbox_events']} through the peer-mapped mailboxes, {c['persistent_chunks']} persistent chunks", flush=True)
dist.barrier()
dist.destroy_process_group()
