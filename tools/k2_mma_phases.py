"""Per-warp phase clocks of the tensor-core K2 variant (debug build: make -C camera-intrinsic-calibration-rs_b200/csrc timing).
usage: CCRS_B200_LIB=camera-intrinsic-calibration-rs_b200/libccrs_b200_timing.so python tools/k2_mma_phases.py [model] [frames]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
model = sys.argv[1] if len(sys.argv) > 1 else "kb4"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
s = c.synth.make_calib(model, nf, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
lib = c._abi.load()
for flush in (True, False):
    ms = gp.time_linearize(s.init_params, reps=5, flush_l2=flush)
    buf = np.zeros((40000, 12), dtype=np.int64)
    lib.ccrs_debug_k2_timing.restype = C.c_int
    nw = lib.ccrs_debug_k2_timing(gp.h, buf.ctypes.data_as(C.c_void_p), 40000)
    b = buf[:nw]
    b = b[b[:, 11] > 0]
    life = b[:, 7] - b[:, 2]
    print(f"flush_l2={flush}: K2 {ms*1e3:.2f} us, {len(b)} warps with frames; span {(b[:, 8].max() - b[:, 0].min())/1e3:.2f} us (globaltimer)")
    print(f"  launch skew ns: median {np.median(b[:,0]-b[:,0].min()):.0f} max {(b[:,0]-b[:,0].min()).max()}")
    print(f"  warp lifetime cycles: median {np.median(life):.0f} min {life.min()} max {life.max()}")
    for name, col in (("setup", None), ("prologues", 4), ("main loops", 5), ("epilogues", 6)):
        v = (b[:, 3] - b[:, 2]) if col is None else b[:, col]
        print(f"  {name:11s} median {np.median(v):8.0f}  p90 {np.percentile(v, 90):8.0f}  max {v.max():8d}")
    print(f"  frames per warp: {np.bincount(b[:, 11].astype(int)).tolist()}; iterations per warp median {np.median(b[:,10]):.0f}; "
          f"loop cycles per iteration median {np.median(b[:,5]/np.maximum(b[:,10],1)):.0f}")
    per_sm = np.bincount(b[:, 1].astype(int), minlength=148)
    print(f"  warps per SM: min {per_sm.min()} max {per_sm.max()}")
gp.close()
