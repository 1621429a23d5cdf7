#!/bin/bash
# E (tests, prefetch A/B, bench) then D (ncu captures) in one call
bash tools/gpu_r2_e.sh
bash tools/gpu_r2_d.sh
