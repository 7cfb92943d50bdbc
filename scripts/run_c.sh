#!/bin/bash
MODELS=relxillCp,relxilllpCp N=6 python scripts/gpu_check.py > gpurun_out/check_cp.log 2>&1; grep -E "final|==|Error|error" gpurun_out/check_cp.log | cut -c1-400 | head -30
NZ=50 MODELS=relxilllpCp N=4 STAGES=0 python scripts/gpu_check.py 2>&1 | grep -E "final|==|rror" | cut -c1-200
