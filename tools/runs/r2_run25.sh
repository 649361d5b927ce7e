#!/bin/bash
# compute-sanitizer memcheck over the new kernels (ingest, fp32 sweep, packed decode, concurrent sweeps) on small tests
T=${1:-r3e}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 3 --launch-timeout 0 python -m pytest -q -x -m gpu \
   tests/test_gpu_ingest.py "tests/test_gpu_f32.py::test_fp32_sweep_against_oracle_every_plane_count" \
   "tests/test_gpu_f32.py::test_fp32_normaliser_underflow_is_redone_in_log_space" \
   "tests/test_gpu_lanes.py::test_stream_encoding_follows_the_counts" \
   "tests/test_gpu_lanes.py::test_underflowed_nonzeros_are_queued_and_redone_in_log_space" \
   "tests/test_gpu_engine.py::test_edge_shapes" "tests/test_gpu_minibatch.py::test_minibatch_default_order_matches_reference" \
   > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/${T}_memcheck.log | tail -12
timeout 600 $CS --tool racecheck --error-exitcode 3 python -m pytest -q -x -m gpu \
   "tests/test_gpu_f32.py::test_fp32_sweep_against_oracle_every_plane_count" \
   "tests/test_gpu_lanes.py::test_stream_encoding_follows_the_counts" > gpurun_out/${T}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${T}_racecheck.log | tail -8
