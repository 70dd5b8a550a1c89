"""Generates tests/golden/golden_runs.npz: small seeded runs of the CPU oracle (inputs, replay
tape, outputs).  The reference itself is Julia and cannot run in this image, so these vectors pin
the ORACLE (against drift) and give the GPU tests a committed fixture; they are not outputs of the
reference.  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import common  # noqa: E402
from common import O, make_case  # noqa: E402

CASES = {
    "gaussian": dict(G=3, Np=4, n_iter=6, kw=dict(burnin=3, theta_snooker=0.25, alpha=0.5, kappa=0.9)),
    "mvnormal": dict(G=2, Np=5, n_iter=5, kw=dict(burnin=2, theta_snooker=0.2, alpha=0.5)),
    "lba": dict(G=2, Np=4, n_iter=5, kw=dict(burnin=2, alpha=0.5, beta=0.3)),
}


def main():
    out = {}
    for i, (name, c) in enumerate(CASES.items()):
        case = make_case(name, np.random.default_rng(100 + i), n_obs=40)
        theta0 = case.theta0(np.random.default_rng(200 + i), c["G"] * c["Np"])
        cfg = case.oracle_config(c["G"], c["Np"], seed=300 + i, **c["kw"])
        r = O.run(cfg, case.oracle_model(), theta0, c["n_iter"])
        out[f"{name}/theta0"] = theta0
        for k, v in case.data.items():
            out[f"{name}/data/{k}"] = np.asarray(v)
        for k, v in r["tape"].items():
            if v is not None:
                out[f"{name}/tape/{k}"] = v
        for k in ("samples", "accept", "lp", "final_id", "final_theta", "final_weight"):
            out[f"{name}/out/{k}"] = r[k]
        out[f"{name}/out/prop_theta"] = r["trace"]["prop_theta"]
        out[f"{name}/out/prop_weight"] = r["trace"]["prop_weight"]
    np.savez_compressed(os.path.join(HERE, "golden_runs.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
