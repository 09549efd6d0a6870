"""Host-side phase breakdown of the LM iteration on the bench workload (ccrs_step_trace)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
s = c.synth.make_calib("eucm", nf, seed=3)
gp = c.Problem.from_synth(s)
lib = c._abi.load()
names = ["K3 launch call", "K3 exec + publish", "host solve", "K2 launch call", "K2 exec + publish", "host accept/reject"]
for flush in (True, False):
    gp.bench_lm_steps(s.init_params, s.init_poses, warmup=3, steps=4, flush_l2=flush)
    lib.ccrs_step_trace(1, None, None)
    ms, _ = gp.bench_lm_steps(s.init_params, s.init_poses, warmup=3, steps=40, flush_l2=flush)
    out = (C.c_double * 6)(); n = C.c_int64(0)
    lib.ccrs_step_trace(0, out, C.byref(n))
    print(f"flush_l2={flush}: {np.mean(ms)*1e3:.1f} us per LM iteration (events), {n.value} traced iterations")
    for nm, v in zip(names, out):
        print(f"   {nm:22s} {v:7.2f} us")
    print(f"   sum                    {sum(out):7.2f} us")
gp.close()
