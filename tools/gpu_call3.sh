#!/bin/bash
set -x
mkdir -p gpurun_out/r02 gpurun_out/ab
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02/pytest_gpu3.log 2>&1
tail -6 gpurun_out/r02/pytest_gpu3.log
bash tools/ab.sh base dopri pref main 2>&1 | tee gpurun_out/ab/summary.txt
