"""BASELINE configs[4] from ONE process: MVN d=100, 1e5 observations, 64 groups x 4096 particles over the 8 GPUs of the
box through a multi-device handle (demcmc_config.n_devices: one host thread per GPU, migration through the peer-mapped
mailboxes), with thinning (store_every) so the history stays bounded, and the posterior summary -- pooled moments,
split-R-hat, ESS -- computed on the devices without downloading the draws (VERDICT r01 items 6 + 7).

  python scripts/c5_thinned_demo.py [n_gpus=8] [n_iter=2000] [store_every=10] [groups_per_gpu=8] [Np=4096]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demcmc_b200 as D  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 10
gpg = int(sys.argv[4]) if len(sys.argv) > 4 else 8
Np = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
dm, n_obs = 100, 100_000
D._ffi.use_library(D._ffi.DEFAULT_LIB)
rng = np.random.default_rng(50514)
mu = rng.normal(size=dm)
x = rng.normal(mu, 1.0, size=(n_obs, dm))
G = gpg * n_gpus
P = G * Np
prior = [("normal", 0.0, 1.0)] * dm + [("halfcauchy", 0.0, 1.0)]
lo, hi = [-np.inf] * dm + [0.0], [np.inf] * (dm + 1)
xbar, s_pool = x.mean(axis=0), float(np.sqrt(((x - x.mean(axis=0)) ** 2).sum() / x.size))
theta0 = np.column_stack([xbar + rng.normal(0, s_pool / np.sqrt(n_obs), size=(P, dm)), s_pool * (1 + rng.normal(0, 5e-4, P))])
t0 = time.perf_counter()
with D.Handle(G, Np, dm + 1, lo, hi, burnin=0, theta_snooker=0.1, seed=20261017, devices=list(range(n_gpus)), store_every=every) as h:
    h.set_model("mvnormal", prior, x=x)
    h.set_state(theta0)
    t1 = time.perf_counter()
    h.run(n_iter)
    t2 = time.perf_counter()
    c = h.counters()
    rows = h.n_rows
    keep0 = rows // 2
    cnt, mean, var = h.moments(keep0, rows - keep0)
    t3 = time.perf_counter()
    rhat, ess = h.diagnostics(keep0, rows - keep0)
    t4 = time.perf_counter()
z = (mean[:dm] - n_obs * xbar / (n_obs + 1.0)) / (s_pool / np.sqrt(n_obs + 1.0))
out = {"what": f"configs[4] shape from one process: MVN d={dm}, {n_obs} obs, {G} groups x {Np} particles on {n_gpus} GPU(s), {n_iter} iterations, store_every={every}",
       "particle_updates_per_s_wall": P * n_iter / (t2 - t1), "particle_updates_per_s_device": c["particle_updates"] / (c["device_ms"] * 1e-3),
       "seconds": {"setup": t1 - t0, "run": t2 - t1, "moments": t3 - t2, "diagnostics": t4 - t3},
       "stored_rows": rows, "history_bytes_per_gpu": rows * (P // n_gpus) * (dm + 1) * 8, "history_bytes_per_gpu_unthinned": n_iter * (P // n_gpus) * (dm + 1) * 8,
       "cross_device_migrations": c["cross_migrations"], "through_mailboxes": c["mailbox_events"], "kernel_launches": c["kernel_launches"],
       "posterior": {"draws_pooled": int(cnt), "max_abs_z_of_pooled_means": float(np.max(np.abs(z))), "sd_ratio_min_max": [float(np.sqrt(var[:dm]).min() / (s_pool / np.sqrt(n_obs + 1))), float(np.sqrt(var[:dm]).max() / (s_pool / np.sqrt(n_obs + 1)))],
                     "max_split_rhat": float(np.nanmax(rhat)), "min_ess": float(np.nanmin(ess)), "rows_used": rows - keep0}}
print(json.dumps(out), flush=True)
