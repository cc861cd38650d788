#!/bin/bash
# round 2, GPU call 30: second derivatives (n = 2) - golden blocks, finite differences; first-derivative tests as regression
set -u
D=gpurun_out/r2c30; mkdir -p $D
( timeout 900 python -m pytest tests -m gpu -q -x -k "derivative" -s ) > $D/pytest_deriv.log 2>&1
tail -25 $D/pytest_deriv.log
