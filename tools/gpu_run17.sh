#!/bin/bash
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu17.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu17.log
