"""Where sample() spends its time outside the kernels (configs[1], 20 iterations): python scripts/e2e_parts.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import demcmc_b200 as D
from demcmc_b200 import api
D._ffi.use_library(D._ffi.DEFAULT_LIB)
x, prior, lo, hi, theta0 = bench.workload(4)
xh = torch.from_numpy(x).pin_memory().numpy()
rng = np.random.default_rng(7)
model = D.DEModel(sample_prior=lambda: [rng.normal(size=50), abs(rng.standard_cauchy())], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                  loglike=D.GPULoglike("mvnormal", xh), names=("μ", "σ"))
de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=4, Np=256, burnin=0, θsnooker=0.1, seed=11)
D.sample(model, de, 20)
import cProfile, pstats
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
D.sample(model, de, 20)
pr.disable()
print("sample(20): %.2f ms under cProfile" % ((time.perf_counter() - t0) * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
