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
ticks = np.zeros((B, 16), dtype=np.int64)
pkg.capi.check(pkg.capi.lib().cmpc_debug_phase_ticks(nc._h, pkg.capi.ptr(ticks)))
names=["load","rk4","powers","E-gemm","scan","conv","gram","reduce+store"]
tt=ticks[:, :8]
dd=np.diff(np.concatenate([np.zeros((B,1),dtype=np.int64), tt], axis=1), axis=1)
for n_,v,m in zip(names, dd.mean(0), np.median(dd,0)): print(f"  {n_:14s} mean {v:9.0f} median {m:9.0f}")
print("stages0-2", (ticks[:,9]-ticks[:,1]).mean(), "horner", (ticks[:,8]-ticks[:,9]).mean(), "stages3+", (ticks[:,2]-ticks[:,8]).mean())
print("total", tt[:,-1].mean(), " | powers: up to end of Horner (stage 4 start):", (ticks[:,8]-ticks[:,0]).mean())
