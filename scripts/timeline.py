"""Kernel timeline of the level chain from %globaltimer stamps (no events, no profiler).
Usage on the GPU box:  DEMCMC_LANES=1 python scripts/timeline.py [out.csv]   -> prints per-level gaps"""
import os, sys, csv
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
os.environ["DEMCMC_TIMELINE"] = "4000"
os.environ["DEMCMC_TIMELINE_FILE"] = out
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
x, prior, lo, hi, theta0 = bench.workload(4)
h = D.Handle(4, 256, 51, lo, hi, burnin=0, theta_snooker=0.1, seed=20261017)
h.set_model("mvnormal", prior, x=x)
h.set_state(theta0)
h.run(60)
h.close()
rows = list(csv.DictReader(open(out)))
R = [{k: float(v) for k, v in r.items()} for r in rows]
R = [r for r in R if r["n"] > 0][40:]                      # skip the warm-up levels
def mean(f): return float(np.mean([f(r) for r in R]))
print("levels", len(R), "mean n", mean(lambda r: r["n"]))
print("propose            %.2f us" % mean(lambda r: r["propose_end"] - r["propose_start"]))
print("propose -> xdot    %.2f us (first xdot CTA start after propose end)" % mean(lambda r: r["xdot_start"] - r["propose_end"]))
print("xdot CTA start spread %.2f us" % mean(lambda r: r["xdot_last_start"] - r["xdot_start"]))
print("xdot start -> wait done %.2f us" % mean(lambda r: r["xdot_wait_done"] - r["xdot_start"]))
print("propose end -> xdot wait done %.2f us" % mean(lambda r: r["xdot_wait_done"] - r["propose_end"]))
print("wait done -> first data %.2f us" % mean(lambda r: r["xdot_first_data"] - r["xdot_wait_done"]))
print("xdot main loop (wait done -> last loop end) %.2f us" % mean(lambda r: r["xdot_loop_end_max"] - r["xdot_wait_done"]))
print("loop end spread (max - min) %.2f us" % mean(lambda r: r["xdot_loop_end_max"] - r["xdot_loop_end_min"]))
print("xdot tail (loop end -> end) %.2f us" % mean(lambda r: r["xdot_end"] - r["xdot_loop_end_max"]))
print("xdot total %.2f us" % mean(lambda r: r["xdot_end"] - r["xdot_start"]))
print("xdot end -> accept start %.2f us" % mean(lambda r: r["accept_start"] - r["xdot_end"]))
print("accept             %.2f us" % mean(lambda r: r["accept_end"] - r["accept_start"]))
nxt = [R[i + 1]["propose_start"] - R[i]["accept_end"] for i in range(len(R) - 1)]
print("accept end -> next propose start %.2f us" % float(np.mean(nxt)))
print("level period %.2f us" % float(np.mean([R[i + 1]["propose_start"] - R[i]["propose_start"] for i in range(len(R) - 1)])))
ideal = mean(lambda r: np.ceil(r["n"] / 8) * 1563 * 8 * 13 * 16 / (148 * 4) / 1965.0)
print("ideal DMMA time of the level %.2f us" % ideal)
cta_file = out + ".cta"
if os.path.exists(cta_file):
    C = [{k: float(v) for k, v in r.items()} for r in csv.DictReader(open(cta_file))]
    dur = np.array([c["loop_end"] - c["wait_done"] for c in C]); tiles = np.array([c["tiles"] // 10 for c in C]); noct = np.array([c["tiles"] % 10 for c in C])
    print("per-CTA loop durations of one level: n_cta %d, min %.1f median %.1f p90 %.1f max %.1f us; wait_done spread %.1f us" % (
        len(C), dur.min(), np.median(dur), np.percentile(dur, 90), dur.max(), max(c["wait_done"] for c in C) - min(c["wait_done"] for c in C)))
    per_tile = dur / (tiles * noct / 4.0)
    print("us per (tile x 4 octets): min %.3f median %.3f max %.3f; ideal at 2 CTAs/SM %.3f" % (per_tile.min(), np.median(per_tile), per_tile.max(), 13 * 8 * 16 * 2 / 1965.0))
    for q in (0, len(C) // 4, len(C) // 2, 3 * len(C) // 4, len(C) - 1):
        c = sorted(C, key=lambda c: c["loop_end"])[q]
        print("  cta %4d tiles %3d oct %d start %.1f wait %.1f end %.1f" % (c["cta"], c["tiles"] // 10, c["tiles"] % 10, c["start"], c["wait_done"], c["loop_end"]))
