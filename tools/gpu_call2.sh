#!/bin/bash
# GPU call 2: whole parity suite (incl. general configurations, timing window), then racecheck on the smallest runs.
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
( time python -m pytest tests -m gpu -q ) > $O/pytest_gpu2.log 2>&1
tail -15 $O/pytest_gpu2.log
python tools/san.py > $O/san_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/san.py > $O/racecheck.log 2>&1
tail -12 $O/racecheck.log
