#!/bin/bash
# static count of local-memory (spill) instructions per kernel in an object file: scripts/spills.sh relxill_b200/build/conv.cu.o
cuobjdump -sass "$1" | awk '/Function :/{f=$3} /STL|LDL/{n[f]++} /^\s+\/\*[0-9a-f]+\*\/ /{t[f]++} END{for(k in t) printf "%s spill_instrs=%d total=%d\n", k, n[k], t[k]}'
