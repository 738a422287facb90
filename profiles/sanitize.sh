#!/bin/bash
# memcheck + racecheck of every kernel family (profiles/sanitize_target.py); logs are committed under profiles/
compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_target.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r02.log
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 3 python profiles/sanitize_target.py > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r02.log
tail -5 gpurun_out/sanitizer_memcheck_r02.log gpurun_out/sanitizer_racecheck_r02.log
