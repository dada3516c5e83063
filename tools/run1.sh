# scratch: one-off GPU checks (overwritten as needed)
python -m pytest tests/test_gpu_parity.py -x -q -k "closed_loop or observer or pieces or streaming" 2>&1 | tail -3
bash tools/ab.sh main
CMPC_B200_LIB=$PWD/build/lib_ticks.so python tools/ticks.py 2>&1 | grep -E "advance_kernel|K3 last block done|K3 first block past"
