#!/bin/bash
# A/B measurement of experiment builds (build/lib_<name>.so, see compressor-mpc_b200/Makefile `variant`) against the
# shipped library: the headline loop and a short sweep per variant, one summary line each.
#   bash tools/ab.sh name1 name2 ...      ("main" = compressor-mpc_b200/libcmpc_b200.so)
mkdir -p gpurun_out/ab
for v in "$@"; do
  if [ "$v" = main ]; then lib=compressor-mpc_b200/libcmpc_b200.so; else lib=build/lib_$v.so; fi
  for rep in $(seq 1 ${REPS:-2}); do
    CMPC_B200_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-b1 --sweep-steps 40 \
        > gpurun_out/ab/$v.$rep.json 2> gpurun_out/ab/$v.$rep.err || tail -3 gpurun_out/ab/$v.$rep.err
    python - "$v" "$rep" <<'PY'
import json, sys
v, rep = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/ab/{v}.{rep}.json"))
except Exception as e:
    print(v, rep, "no result", e); raise SystemExit
r, s = d["roofline"], d["sweep"]
print(f"{v:10s} #{rep} value {d['value']/1e6:6.2f} M  ms/step {d['ms_per_step']*1e3:6.1f} us  b2b {d['value_back_to_back']/1e6:6.2f} M  "
      f"asm {r['kernel_ms']*1e3:6.1f} us ({r['frac']:.3f})  ctrl {r['control_step_ms']*1e3:6.1f} us  e2e {d['e2e']['value']/1e6:5.2f} M (pipelined {d.get('e2e_pipelined',{}).get('value',0)/1e6:5.2f} M)  "
      f"ctl-e2e {d['e2e_control_step']['value']/1e6:5.2f} M | sweep {s['value']/1e6:6.2f} M asm {s['roofline']['kernel_ms']*1e3:7.1f} us "
      f"({s['roofline']['frac']:.3f}) parity {d['health']['gpu_vs_oracle_max_rel_err_u']:.1e}/{s['health']['gpu_vs_oracle_max_rel_err_u']:.1e} "
      f"act {d['health']['active_sets_identical']}/{s['health']['active_sets_identical']} fail {d['health']['qp_failures']}+{s['health']['qp_failures']}")
PY
  done
done
