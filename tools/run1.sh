python -m pytest tests/test_gpu_parity.py -x -q -k "page_locked or streaming" 2>&1 | tail -3
for v in "" 1; do
  echo "== CMPC_NO_EARLY_RETURN=$v"
  if [ -n "$v" ]; then export CMPC_NO_EARLY_RETURN=1; fi
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-b1 --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.2f M  ms/step %.1f us  e2e %.2f M (%.1f us/record)  diff vs device run %s' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, 4096/d['e2e']['value']*1e6, d['e2e'].get('max_abs_diff_vs_device_run')))"
done
