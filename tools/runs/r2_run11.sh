#!/bin/bash
T=${1:-r2k}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/${T}_tests.log
timeout 600 python tools/time_ingest.py 20000 20000 1000 > gpurun_out/${T}_ingest.json 2> gpurun_out/${T}_ingest.err
echo "ingest timing rc=$?"; cat gpurun_out/${T}_ingest.json; tail -3 gpurun_out/${T}_ingest.err
