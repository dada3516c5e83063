# scratch: one-off GPU checks (overwritten as needed)
python -m pytest tests/test_gpu_parity.py -x -q -k "page_locked or streaming" 2>&1 | tail -3
bash tools/ab.sh main
