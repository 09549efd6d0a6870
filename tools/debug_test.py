import os, sys, faulthandler, subprocess, json
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ccrs_b200 as c
import ctypes as C
from test_gpu_parity import _SOLVE_CODE
s = c.synth.make_calib("eucm", 100, seed=1, noise_px=0.1)
gp = c.Problem.from_synth(s)
lib = c._abi.load()
for loop in ("solve_lm", "solve_gn"):
    gp.set_poses(s.init_poses)
    print("in-process", loop, flush=True)
    intr, summ, hist = getattr(gp, loop)(s.init_params)
    print(" ->", summ.iterations, summ.status, flush=True)
    print("subprocess", loop, flush=True)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _SOLVE_CODE % loop], cwd=root, env=dict(os.environ, CCRS_DEVICE_LOOP="0"), capture_output=True, text=True, timeout=30)
    print(" -> rc", r.returncode, r.stdout[-200:], r.stderr[-500:], flush=True)
gp.close()
print("done", flush=True)
