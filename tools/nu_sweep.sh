#!/bin/bash
# development aid: multigrid smoothing sweeps vs time-to-solution at 1M elements
for nu in 3 2 1; do
  python tools/solve_scaling.py 1048576 adr,diffusion 2 "{\"mg_nu\": $nu}" 2>/dev/null | sed "s/^/nu=$nu /"
done
