#!/bin/bash
# The tree / sort / leapfrog GPU tests against a build with device-side index asserts (-DB200_BOUNDS_CHECK):
# scatter destinations, walk-record ids and successor monotonicity, stored-particle ranges.  A failed assert aborts
# the kernel and surfaces as a CUDA error in the next call.  compute-sanitizer is closed on this pool; this is the
# in-tree substitute.   usage (on a GPU box): bash tools/bounds_check.sh
set -u
cd "$(dirname "$0")/.."
make -s -j8 -C lambda-cdm-raytracing_b200/csrc check-build || exit 1
B200GRAV_LIB=$PWD/lambda-cdm-raytracing_b200/lib/libb200grav_check.so \
  python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_tree_fixed.py tests/test_gpu_leapfrog.py -m gpu -x -q
