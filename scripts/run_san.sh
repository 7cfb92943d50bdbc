#!/bin/bash
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize.py 2>&1 | tail -25 ) > gpurun_out/sanitizer_memcheck.txt
( timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize.py 2>&1 | tail -25 ) > gpurun_out/sanitizer_racecheck.txt
tail -4 gpurun_out/sanitizer_memcheck.txt; tail -4 gpurun_out/sanitizer_racecheck.txt
