#!/bin/bash
# compute-sanitizer memcheck on the kernels added after the previous capture (two legs, GaussianReference, leapfrog entry point).
# Single-handle cases only: several handles on one device need their kernels to run side by side, which the sanitizer's
# serialisation does not allow (the hand-shake would time out).
set -x
O=gpurun_out/r2san2
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_legs_gmm6_automala_gaussian or two_legs_funnel8_slice_gaussian or two_legs_ising5 or two_legs_test_swapper or one_leg_variational or two_legs_funnel40_mala or integrator and toy_mvn or variational_entry" > $O/compute_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?" >> $O/compute_sanitizer_memcheck.log; tail -6 $O/compute_sanitizer_memcheck.log
