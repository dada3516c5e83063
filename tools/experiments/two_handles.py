"""Experiment: the headline batch as ONE handle of 4096 scenarios against TWO handles of 2048 on two streams
(back to back, no L2 flush): does the machine fill the tail / solve / plant phases of one half with the
assemble kernel of the other?    python tools/experiments/two_handles.py [T]"""
import sys, json, pathlib, time
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g
pkg = g.load_package()
s = pkg.setupfile.setup_from_dict(json.load(open(ROOT / "tests/golden/setups.json"))["coop-par"])
x_def, _ = pkg.plant_defaults(0)
B, T = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 200
x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
dev = torch.device("cuda", 0)
def run(parts, prios):
    hs, bufs, streams = [], [], []
    for i, (lo, hi) in enumerate(parts):
        nc = pkg.from_setup(s, batch=hi - lo)
        d = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (x0, be, bo)]
        traj = torch.zeros((hi - lo, T, 20), dtype=torch.float64, device=dev)
        hs.append(nc); bufs.append((d, traj)); streams.append(torch.cuda.Stream(priority=prios[i]))
    def go(first, n):
        for k in range(first, first + n):          # record by record, alternating the handles, so that neither stream runs dry
            for nc, (d, traj), st in zip(hs, bufs, streams):
                nc.run_closed_loop_device(k, 1, T, d[0].data_ptr(), be.shape[1], d[1].data_ptr(), d[2].data_ptr(), traj.data_ptr(), 0, 0, 0, st.cuda_stream)
    go(0, 20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter()
    go(20, T - 20)
    for st in streams: torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return B * (T - 20) / (ms * 1e-3), ms / (T - 20) * 1e3, torch.cat([b[1] for b in bufs]).cpu().numpy()
v1, us1, t1 = run([(0, B)], [0])
print(f"one handle  x 4096: {v1/1e6:6.2f} M steps/s  {us1:6.1f} us per record")
for parts, pr in (([(0, 2048), (2048, 4096)], [0, 0]), ([(0, 2048), (2048, 4096)], [-1, 0]), ([(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)], [0, 0, 0, 0])):
    v2, us2, t2 = run(parts, pr)
    print(f"{len(parts)} handles, stream priorities {pr}: {v2/1e6:6.2f} M steps/s  {us2:6.1f} us per record   identical records: {np.array_equal(t1, t2)}")
