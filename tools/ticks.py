import sys, json, numpy as np, ctypes as C
import pathlib; ROOT = str(pathlib.Path(__file__).resolve().parent.parent); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import __graft_entry__ as g
pkg = g.load_package()
s = pkg.setupfile.setup_from_dict(json.load(open(ROOT + "/tests/golden/setups.json"))["coop-par"])
x_def, u_def = pkg.plant_defaults(0)
B, T = 4096, 30
x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
nc = pkg.from_setup(s, batch=B)
nc.run_closed_loop(x0, be, bo, T, want_traj=False, want_qp=False)
ticks = np.zeros((B, 32), dtype=np.int64)
pkg.capi.check(pkg.capi.lib().cmpc_debug_phase_ticks(nc._h, pkg.capi.ptr(ticks)))
names=["load","rk4","powers","E-gemm","scan","conv","gram","reduce+store"]
tt=ticks[:, :8]
dd=np.diff(np.concatenate([np.zeros((B,1),dtype=np.int64), tt], axis=1), axis=1)
for n_,v,m in zip(names, dd.mean(0), np.median(dd,0)): print(f"  {n_:14s} mean {v:9.0f} median {m:9.0f}")
print("stages0-2", (ticks[:,9]-ticks[:,1]).mean(), "horner", (ticks[:,8]-ticks[:,9]).mean(), "stages3+", (ticks[:,2]-ticks[:,8]).mean())
print("total", tt[:,-1].mean(), " | powers: up to end of Horner (stage 4 start):", (ticks[:,8]-ticks[:,0]).mean())

acc = ticks[:, 14] >> 40
t14 = ticks[:, 14] & ((1 << 40) - 1)
print("solve_kernel  (cycles after its QP data arrived, lane 0 of each scenario): inverse done %.0f, first sweep done %.0f, sweeps done %.0f, end %.0f" % (ticks[:,30].mean(), ticks[:,31].mean(), ticks[:,11].mean(), ticks[:,12].mean()))
print("integrator (same clock): stage 0 starts %.0f, stage 1 starts %.0f, stage 2 starts %.0f, stages done %.0f, error norm done %.0f" % tuple(ticks[:, i].mean() for i in (25, 26, 27, 28, 29)))
print("advance_kernel (cycles after its wait): inputs arrived %.0f, integration done %.0f, linearisation stored %.0f; accepted steps: %s" % (ticks[:,13].mean(), t14.mean(), ticks[:,15].mean(), np.bincount(acc.astype(int)).tolist()))

g = ticks[:, 16:].astype(np.float64)
def mm(col, f):
    v = g[:, col - 16]; v = v[v > 0]
    return f(v)
t0 = mm(17, np.min)
ev = [("K1 first CTA resident", mm(16, np.min)), ("K1 first CTA past its wait", t0), ("K1 last CTA resident", mm(16, np.max)), ("K1 last CTA done", mm(18, np.max)),
      ("K2 first warp resident", mm(19, np.min)), ("K2 last warp resident", mm(19, np.max)), ("K2 first warp past its wait", mm(20, np.min)), ("K2 last warp past its wait", mm(20, np.max)), ("K2 last warp done", mm(21, np.max)),
      ("K3 first block resident", mm(22, np.min)), ("K3 first block past its wait", mm(23, np.min)), ("K3 last block past its wait", mm(23, np.max)), ("K3 last block done", mm(24, np.max))]
print("timeline of the last record (us after the first assemble CTA passed its wait; %globaltimer):")
for n_, v in ev: print(f"  {n_:30s} {(v - t0) / 1e3:8.2f}")

tot = tt[:, -1].astype(np.float64)
print("per-scenario assemble time (cycles): p5 %.0f p50 %.0f p95 %.0f max %.0f" % tuple(np.percentile(tot, [5, 50, 95, 100])))
import os
W = int(os.environ.get("CMPC_TICKS_WAVE", "592"))
cta = np.arange(B) % W
per_cta = np.bincount(cta, weights=tot, minlength=W)
print("per-CTA sum over its scenarios (persistent grid of %d): min %.0f mean %.0f max %.0f cycles = %.1f / %.1f / %.1f us" % (W, per_cta.min(), per_cta.mean(), per_cta.max(), per_cta.min()/1965, per_cta.mean()/1965, per_cta.max()/1965))
rnd = np.arange(B) // W
for r in range(rnd.max() + 1): print("  round %d: mean %.0f max %.0f" % (r, tot[rnd == r].mean(), tot[rnd == r].max()))
e18 = g[:, 18 - 16]
print("assemble CTA end times (us after start): p5 %.1f p50 %.1f p95 %.1f max %.1f (last round only)" % tuple((np.percentile(e18[rnd == rnd.max()], [5, 50, 95, 100]) - t0) / 1e3))

if os.environ.get("CMPC_TICKS_SMID"):
    smid = ticks[:, 19]
    sms = np.unique(smid)
    m = np.array([tot[smid == k].mean() for k in sms]); n = np.array([(smid == k).sum() for k in sms])
    o = np.argsort(m)
    print("SMs seen", len(sms), " scenarios per SM: min %d max %d" % (n.min(), n.max()))
    print("per-SM mean scenario time, fastest 10:", [(int(sms[i]), int(m[i]), int(n[i])) for i in o[:10]])
    print("slowest 10:", [(int(sms[i]), int(m[i]), int(n[i])) for i in o[-10:]])
    print("by smid parity: even %.0f odd %.0f" % (m[sms % 2 == 0].mean(), m[sms % 2 == 1].mean()))
    for lo in range(0, 160, 16): 
        sel = (sms >= lo) & (sms < lo + 16)
        if sel.any(): print("  smid %3d..%3d: mean %.0f  scenarios %d" % (lo, lo + 15, m[sel].mean(), n[sel].sum()))
