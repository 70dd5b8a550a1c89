import sys, os, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import demcmc_b200 as D
lib = sys.argv[1]
D._ffi.DEFAULT_LIB = lib
sys.argv = ["bench_configs.py", "c4"]
exec(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench_configs.py")).read())
