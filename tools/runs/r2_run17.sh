#!/bin/bash
T=${1:-r2r}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_minibatch.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
timeout 600 python tools/time_minibatch.py > gpurun_out/${T}_minibatch_timing.json 2> gpurun_out/${T}_minibatch_timing.err
echo "minibatch timing rc=$?"; cat gpurun_out/${T}_minibatch_timing.json; tail -3 gpurun_out/${T}_minibatch_timing.err
