#!/bin/bash
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --print-limit 20 python tools/prof_one.py --shape 4096x1024 --effort 0.25 --n 3 --copies 1 > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
tail -n 6 gpurun_out/r02_sanitizer_memcheck.txt
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python tools/prof_one.py --shape 4096x1024 --effort 0.25 --n 2 --copies 1 > gpurun_out/r02_sanitizer_racecheck.txt 2>&1
tail -n 6 gpurun_out/r02_sanitizer_racecheck.txt
