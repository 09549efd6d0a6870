"""Per-warp phase clocks of K2 (debug build: make -C camera-intrinsic-calibration-rs_b200/csrc timing).

phases: 0 start | 1 prologue done | 2 main loop done | 3 basis change done | 4 reduction+store done | 5 end
usage: CCRS_B200_LIB=camera-intrinsic-calibration-rs_b200/libccrs_b200_timing.so python tools/k2_phases.py [model] [frames]
"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
model = sys.argv[1] if len(sys.argv) > 1 else "eucm"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
s = c.synth.make_calib(model, nf, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
lib = c._abi.load()
for flush in (True, False):
    ms = gp.time_linearize(s.init_params, reps=5, flush_l2=flush)
    buf = np.zeros((20000, 12), dtype=np.int64)
    lib.ccrs_debug_k2_timing.restype = C.c_int
    nw = lib.ccrs_debug_k2_timing(gp.h, buf.ctypes.data_as(C.c_void_p), 20000)
    b = buf[:nw]
    gt = b[:, 0] - b[:, 0].min()
    ph = np.diff(b[:, 2:8], axis=1)
    names = ["prologue", "main loop", "basis chg", "reduce+store", "stats"]
    print(f"flush_l2={flush}: K2 {ms*1e3:.2f} us, {nw} warps; launch skew (globaltimer ns) median {np.median(gt):.0f} max {gt.max()}")
    print(f"  of reduce+store: cost shuffles + partial publish + release ticket: median {np.median(b[:, 10] - b[:, 5]):.0f} cycles")
    tot = (b[:, 7] - b[:, 2])
    print(f"  warp lifetime cycles: median {np.median(tot):.0f} max {tot.max()} min {tot.min()}")
    for i, n in enumerate(names):
        print(f"  {n:13s} median {np.median(ph[:, i]):8.0f}  p90 {np.percentile(ph[:, i], 90):8.0f}  max {ph[:, i].max():8d}")
    span_ns = b[:, 8].max() - b[:, 0].min()
    mhz = np.median(tot / np.maximum(b[:, 8] - b[:, 0], 1)) * 1e3
    print(f"  first warp start -> last warp end: {span_ns/1e3:.2f} us (globaltimer); SM clock from clock64/globaltimer: {mhz:.0f} MHz")
    ml = ph[:, 1]
    print("  main-loop histogram (kcycles):", np.histogram(ml / 1e3, bins=8)[0].tolist(), [round(x, 1) for x in np.histogram(ml / 1e3, bins=8)[1].tolist()])
    slow = ml > np.percentile(ml, 85)
    print(f"  slow warps: SMs {np.unique(b[slow, 1]).size}, warp slots {np.bincount(b[slow, 9].astype(int) % 4, minlength=4).tolist()} (hw warpid % 4), all: {np.bincount(b[:, 9].astype(int) % 4, minlength=4).tolist()}")
    persm = np.array([ml[b[:, 1] == i].mean() if np.any(b[:, 1] == i) else 0 for i in range(148)])
    print("  main-loop mean kcycles by smid:", " ".join(f"{x/1e3:.0f}" for x in persm))
    sm = b[:, 1]
    per_sm = np.bincount(sm.astype(int))
    print(f"  warps per SM: min {per_sm[per_sm>0].min()} max {per_sm.max()} SMs used {np.count_nonzero(per_sm)}")
# K3 phases (one reduce with damping)
gp.set_poses(s.init_poses)
gp.linearize(s.init_params)
for rep in range(3):
    gp.reduce(0, 1e-4)
buf = np.zeros((20000, 8), dtype=np.int64)
lib.ccrs_debug_k3_timing.restype = C.c_int
nw = lib.ccrs_debug_k3_timing(gp.h, buf.ctypes.data_as(C.c_void_p), 20000)
b = buf[:nw]
ph = np.diff(b[:, 0:6], axis=1)
print(f"K3: {nw} warps")
for i, n in enumerate(["load+cholesky", "solve+reduce-terms+stores", "CTA reduction", "fence+ticket", "last CTA / tail"]):
    print(f"  {n:26s} median {np.median(ph[:, i]):8.0f}  max {ph[:, i].max():8d}")
print(f"  warp lifetime median {np.median(b[:,5]-b[:,0]):.0f} max {(b[:,5]-b[:,0]).max()}; span first start -> last end {b[:,5].max()-b[:,0].min()} (clock64 differs per SM: indicative)")
gp.close()
