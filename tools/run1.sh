# scratch: one-off GPU checks (overwritten as needed)
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time python bench.py > gpurun_out/r02c/bench_default.json 2> gpurun_out/r02c/bench_default.err ) 2>&1 | tail -3
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02c/bench_default.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','steps','warmup','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'])"
( time python bench.py --impl reference > gpurun_out/r02c/bench_ref_default.json 2>> gpurun_out/r02c/bench_default.err ) 2>&1 | tail -3
tail -1 gpurun_out/r02c/bench_ref_default.json | cut -c1-300
