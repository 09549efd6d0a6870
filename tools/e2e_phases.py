"""Wall-clock phases of one end-to-end calibration call with host buffers (the bench's e2e leg)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
import torch
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
s = c.synth.make_calib("eucm", nf, seed=3)
def pin(a):
    t = torch.empty(a.shape, dtype={np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[a.dtype], pin_memory=True)
    n = t.numpy(); n[...] = a; return n, t
f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
keep = []
arrs = []
for a in (f32(s.x), f32(s.y), f32(s.z), f32(s.u), f32(s.v)):
    n, t = pin(a); keep.append(t); arrs.append(n)
fo, t = pin(s.frame_offsets); keep.append(t)
hid, t = pin(np.ascontiguousarray(s.extra["corner_id"], dtype=np.int32)); keep.append(t)
hboard = np.ascontiguousarray(s.extra["board"], dtype=np.float32)
hout, t = pin(np.zeros_like(s.init_poses)); keep.append(t)
fmt = sys.argv[2] if len(sys.argv) > 2 else "board"
hp, t = pin(np.ascontiguousarray(s.init_poses)); keep.append(t)
acc = np.zeros(6); N = 12
for i in range(N + 2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if fmt == "board":
        q = c.Problem("eucm", s.width, s.height, fo, None, None, None, arrs[3], arrs[4], device=0, corner_id=hid, board=hboard)
    else:
        q = c.Problem("eucm", s.width, s.height, fo, *arrs, device=0)
    t1 = time.perf_counter()
    q.set_poses(hp)
    t2 = time.perf_counter()
    intr, summ, _ = q.solve_lm(s.init_params)
    t3 = time.perf_counter()
    out = q.get_poses(out=hout)
    t4 = time.perf_counter()
    q.close()
    t5 = time.perf_counter()
    if i >= 2:
        acc += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0])
names = [f"create ({fmt} format, H2D obs)", "set_poses (H2D)", f"solve_lm ({summ.iterations} its)", "get_poses (D2H)", "destroy", "total"]
for n, v in zip(names, acc / N):
    print(f"{n:28s} {v*1e6:9.1f} us")
