# scratch: one-off GPU checks (overwritten as needed)
python -m pytest tests/test_gpu_parity.py -x -q -k "control_step or workflow or argument or custom or ragged" 2>&1 | tail -3
bash tools/ab.sh main
CMPC_NO_DIRECT_HOST_IO=1 bash tools/ab.sh main
