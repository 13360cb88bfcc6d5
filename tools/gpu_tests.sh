#!/bin/bash
# The driver's round-end GPU tier in one call: full GPU suite, then the smoke entry point.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
