"""A/B of the speculative K3 launch on a converging LM solve (run twice: CCRS_SPEC_K3=1 / 0)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
s = c.synth.make_calib("eucm", nf, seed=3, noise_px=0.1)
gp = c.Problem.from_synth(s)
lib = c._abi.load()
ms, its = [], 0
for rep in range(25):
    gp.set_poses(s.init_poses)
    intr, summ, _ = gp.solve_lm(s.init_params)
    if rep >= 5:
        ms.append(summ.device_ms); its = summ.iterations
a, h = C.c_int64(0), C.c_int64(0)
en = lib.ccrs_spec_k3_counters(C.byref(a), C.byref(h))
print(f"spec={en} iterations={its} device_ms median={np.median(ms):.4f} min={np.min(ms):.4f} per-iteration={np.median(ms)/its*1e3:.1f} us; spec launches={a.value} hits={h.value}")
gp.close()
