#!/bin/bash
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_target.py > gpurun_out/sanitizer_memcheck_state_4.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_state_4.log; tail -4 gpurun_out/sanitizer_memcheck_state_4.log
