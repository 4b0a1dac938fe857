#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s "$@" > gpurun_out/r02_tests.log 2>&1
grep -E "^\[parity\]|passed|failed|Error|error" gpurun_out/r02_tests.log | tail -40
tail -5 gpurun_out/r02_tests.log
