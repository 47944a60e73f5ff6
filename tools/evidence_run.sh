#!/bin/bash
# Round evidence on one B200: bench line, ncu launch list of the same command, full ncu capture of the dominant kernel,
# reference arm.  Usage (from the repo root, on the GPU box): bash tools/evidence_run.sh r01
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; echo rc_bench=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-solve > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-solve > gpurun_out/ncu_launch.log 2>&1; echo rc_launches=$?
ncu --set full --clock-control none --import-source on -k regex:dg_assemble_elem -s 3 -c 1 -f -o gpurun_out/prof_assemble_${R} \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-solve > gpurun_out/ncu_full.log 2>&1; echo rc_ncu=$?
python bench.py --impl reference > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err; echo rc_ref=$?
