#!/bin/bash
set -x
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu7.log 2>&1
tail -25 gpurun_out/r02/pytest_gpu7.log
python tests/fuzz_parity.py 3 60 general > gpurun_out/r02/fuzz_general.log 2>&1
tail -3 gpurun_out/r02/fuzz_general.log
python tests/fuzz_parity.py 5 60 > gpurun_out/r02/fuzz_tuned.log 2>&1
tail -2 gpurun_out/r02/fuzz_tuned.log
python -c "import __graft_entry__ as g; g.smoke()"
