#!/bin/bash
# round-2 final: the C++ side on one GPU -- plugin parity against the reference's CPU tree, the reference's smoke program
O=gpurun_out
tests/host/_bin/host_parity_test > $O/r2_host_parity.log 2>&1; echo "host_parity rc=$?"
tail -2 $O/r2_host_parity.log
lambda-cdm-raytracing_b200/examples/_bin/nbody_b200 1048576 50 tree zeldovich > $O/r2_nbody_b200_tree.log 2>&1; echo "nbody tree rc=$?"
tail -4 $O/r2_nbody_b200_tree.log
lambda-cdm-raytracing_b200/examples/_bin/nbody_b200 65536 50 direct random > $O/r2_nbody_b200_direct.log 2>&1; echo "nbody direct rc=$?"
tail -3 $O/r2_nbody_b200_direct.log
