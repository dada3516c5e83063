( time timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -x -q -k "runaway" ) 2>&1 | tail -6
