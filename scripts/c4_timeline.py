"""Timeline of the level-by-level path on configs[3] (hierarchical, wide kernels) from in-kernel %globaltimer stamps.
Usage on the GPU box:  python scripts/c4_timeline.py [out.csv] [config]"""
import os, sys, csv
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/c4_timeline.csv"
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
os.environ["DEMCMC_TIMELINE"] = "2000"
os.environ["DEMCMC_TIMELINE_FILE"] = out
os.environ["DEMCMC_HOST_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import bench_configs as BC
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
line = BC.run(name, 12)
print(name, "updates/s %.0f" % line["particle_updates_per_s"], "levels/sweep", line["levels_per_sweep"])
R = [{k: float(v) for k, v in r.items()} for r in csv.DictReader(open(out))]
R = R[40:160]                                            # after the warm-up, before the event-bracketed pass
def mean(f, rows): return float(np.mean([f(r) for r in rows])) if rows else float("nan")
print("levels", len(R), "span %.1f us per level" % ((R[-1]["accept_end"] - R[0]["accept_end"]) / (len(R) - 1)))
bins = [(1, 64), (65, 592), (593, 1184), (1185, 2400), (2401, 10 ** 9)]
for lo, hi in bins:
    S = [(a, b) for a, b in zip(R[:-1], R[1:]) if lo <= b["n"] <= hi]
    if not S: continue
    cur = [b for a, b in S]
    print("n %5d-%-6d levels %3d mean n %6.0f | period %6.1f | prev accept end -> propose wait done %5.1f | wait done -> propose end %5.1f | -> xdot wait done %5.1f | -> xdot end %5.1f | -> accept wait done %5.1f | -> accept end %5.1f" % (
        lo, min(hi, 99999), len(S), mean(lambda r: r["n"], cur), mean(lambda ab: ab[1]["accept_end"] - ab[0]["accept_end"], S),
        mean(lambda ab: ab[1]["propose_wait_done"] - ab[0]["accept_end"], S), mean(lambda r: r["propose_end"] - r["propose_wait_done"], cur),
        mean(lambda r: r["xdot_wait_done"] - r["propose_end"], cur), mean(lambda r: r["xdot_end"] - r["xdot_wait_done"], cur),
        mean(lambda r: r["accept_wait_done"] - r["xdot_end"], cur), mean(lambda r: r["accept_end"] - r["accept_wait_done"], cur)))
    if "q0" in cur[0] and cur[0]["q0"] > 0:
        print("      propose phases (last CTA): wait done -> donors/projection %.1f -> proposal stored %.1f -> priors %.1f -> staged %.1f -> end %.1f" % (
            mean(lambda r: r["q0"] - r["propose_wait_done"], cur), mean(lambda r: r["q1"] - r["q0"], cur), mean(lambda r: r["q2"] - r["q1"], cur),
            mean(lambda r: r["q3"] - r["q2"], cur), mean(lambda r: r["propose_end"] - r["q3"], cur)))
