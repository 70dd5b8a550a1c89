"""torchrun --nproc-per-node N scripts/e2e_parts_dist.py : where distributed.sample() spends its time (configs[1] per rank, 20 iterations)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
import demcmc_b200 as D
from demcmc_b200 import distributed
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D._ffi.use_library(D._ffi.DEFAULT_LIB)
x, prior, lo, hi, theta0 = bench.workload(4 * world)
xh = torch.from_numpy(x).pin_memory().numpy()
def make():
    rng = np.random.default_rng(7)
    model = D.DEModel(sample_prior=lambda: [rng.normal(size=50), abs(rng.standard_cauchy())], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                      loglike=D.GPULoglike("mvnormal", xh), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=4 * world, Np=256, burnin=0, θsnooker=0.1, seed=11)
    return model, de
m, de = make(); distributed.sample(m, de, 16, device=local)
m, de = make(); distributed.sample(m, de, 20, device=local)
import cProfile, pstats
pr = cProfile.Profile()
m, de = make()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
distributed.sample(m, de, 20, device=local)
pr.disable()
dt = time.perf_counter() - t0
if rank == 0:
    print("distributed.sample(20) on %d GPUs: %.2f ms under cProfile" % (world, dt * 1e3))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
dist.barrier(); dist.destroy_process_group()
