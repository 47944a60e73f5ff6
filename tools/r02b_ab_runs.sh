# The A/B runs of round 2, session 3 (one B200 per call: gpurun -- "bash tools/r02b_ab_runs.sh <name>").  Variant libraries come from
# tools/build_variant.sh; results are summarised in profiles/r02_assemble_experiments.md (addendum), r02_solve_launch_summary.txt,
# r02_e2e_*.log and DESIGN.md sections 5-6.
case "$1" in
ab)
  # round 2, session 3: full GPU suite, then A/B of the volume-kernel variants (tools/build_variant.sh) and of the malloc tuning
  set -x
  mkdir -p gpurun_out/r02b
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b/tests.log
  tail -4 gpurun_out/r02b/tests.log
  run() {  # name lib
    if [ -n "$2" ]; then export REYNA_B200_LIB=$PWD/$2; else unset REYNA_B200_LIB; fi
    timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline > gpurun_out/r02b/bench_$1.json 2> gpurun_out/r02b/bench_$1.err
    python -c "
  import json,sys; d=json.load(open('gpurun_out/r02b/bench_$1.json')); print('$1', 'step', round(d['ms_per_step'],4), 'call', round(d['roofline']['kernel_ms'],4), 'e2e_s', round(d['e2e']['s_per_step'],4), d['roofline']['fp64'])"
  }
  run base ""
  for v in pf pf184 r184 lb11; do run $v build/lib_$v.so; done
  run base2 ""
  unset REYNA_B200_LIB
  REYNA_B200_MALLOPT=0 run nomallopt ""
  run mallopt ""
  ;;
solver_ab)
  # round 2, session 3: solver kernels A/B (block-row SpMV, unrolled multigrid block rows, parallel dense inverse) + evidence
  set -x
  mkdir -p gpurun_out/r02b
  lscpu | grep -i "model name\|^CPU(s)\|numa\|thread\|socket" > gpurun_out/r02b/host.txt; nproc >> gpurun_out/r02b/host.txt
  timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_convergence.py -x -q > gpurun_out/r02b/solver_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/solver_tests.log
  tail -3 gpurun_out/r02b/solver_tests.log
  for v in "" build/lib_spmv_old.so build/lib_solv_old.so "" build/lib_solv_old.so; do
    if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
    echo "lib=$v" >> gpurun_out/r02b/solve_ab.log
    timeout 300 python tools/solve_once.py 1048576 adr 2 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab.log
    timeout 300 python tools/solve_once.py 1048576 diffusion 2 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab.log
  done
  unset REYNA_B200_LIB
  cat gpurun_out/r02b/solve_ab.log
  for c in 1 2; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b/bench_config$c.json 2>/dev/null; done
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/solve_launches.csv -c 6000 python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solve.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_blockrow_kernel|mgb_block_row" --launch-skip 30 --launch-count 3 -o gpurun_out/r02b/prof_solver -f python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solver_full.log 2>&1
  ls -la gpurun_out/r02b
  ;;
solver_ab2)
  set -x
  mkdir -p gpurun_out/r02b
  for v in "" build/lib_mgb_late.so build/lib_solv_old.so "" build/lib_mgb_late.so; do
    if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
    timeout 300 python tools/solve_repeat.py 1048576 adr 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
  done
  unset REYNA_B200_LIB
  timeout 300 python tools/solve_repeat.py 1048576 adr 2 4 '{"mg_fp32": true}' 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
  timeout 300 python tools/solve_repeat.py 1048576 diffusion 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
  timeout 300 python tools/solve_repeat.py 65536 adr 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
  timeout 300 python tools/solve_repeat.py 1024 adr 1 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
  cat gpurun_out/r02b/solve_ab2.log
  timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/r02b/solver_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/solver_tests2.log; tail -2 gpurun_out/r02b/solver_tests2.log
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/solve_launches2.csv -c 3000 python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solve2.log 2>&1
  ;;
window)
  set -x
  mkdir -p gpurun_out/r02b
  timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/r02b/solver_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/solver_tests3.log; tail -2 gpurun_out/r02b/solver_tests3.log
  for E in 131072 1048576; do for w in 16 32 64; do
    REYNA_B200_FACE_WINDOW=$w timeout 300 python bench.py --elements $E --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); print('E=$E w=$w', 'step', round(d['ms_per_step'],4), 'call', round(d['roofline']['kernel_ms'],4))" >> gpurun_out/r02b/window.log
  done; done
  cat gpurun_out/r02b/window.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b/bench_n1_solve.json 2>/dev/null; python -c "
  import json; d=json.load(open('gpurun_out/r02b/bench_n1_solve.json')); print(d['solve'])"
  ;;
e2e)
  set -x
  mkdir -p gpurun_out/r02b
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02b/parity_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/parity_tests.log; tail -2 gpurun_out/r02b/parity_tests.log
  for i in 1 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); print('run $i', 'step', round(d['ms_per_step'],4), 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> gpurun_out/r02b/e2e.log
  done
  cat gpurun_out/r02b/e2e.log
  ;;
arena)
  set -x
  O=gpurun_out/r02f
  mkdir -p $O
  timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_probes.py -x -q > $O/solver_tests_arena.log 2>&1; echo "rc=$?" >> $O/solver_tests_arena.log; tail -2 $O/solver_tests_arena.log
  timeout 300 python tools/solve_repeat.py 1048576 adr 2 8 2>/dev/null | tail -1 >> $O/solve_arena.log
  timeout 300 python tools/solve_repeat.py 16384 diffusion 1 8 2>/dev/null | tail -1 >> $O/solve_arena.log
  timeout 300 python tools/solve_repeat.py 1024 adr 1 8 2>/dev/null | tail -1 >> $O/solve_arena.log
  cat $O/solve_arena.log
  for c in 1 2; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config$c.json 2> $O/bench_config$c.err; done
  ;;
minb)
  set -x
  O=gpurun_out/r02f
  mkdir -p $O
  for v in "" build/lib_mgb_minb6.so build/lib_mgb_minb8.so ""; do
    if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
    timeout 300 python tools/solve_repeat.py 1048576 adr 2 5 2>/dev/null | tail -1 >> $O/solve_minb.log
  done
  cat $O/solve_minb.log
  ;;
setup_diag)
  set -x
  O=gpurun_out/r02f; mkdir -p $O
  for i in 1 2 3; do timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('lazy $i', s['dgfem_solve_true_wall_s'], s['solve_s'], s['setup_s'])" >> $O/setup_diag.log; done
  for i in 1 2; do CUDA_MODULE_LOADING=EAGER timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('eager $i', s['dgfem_solve_true_wall_s'], s['solve_s'], s['setup_s'])" >> $O/setup_diag.log; done
  cat $O/setup_diag.log
  ;;
chunk)
  set -x
  O=gpurun_out/r02g; mkdir -p $O
  for rep in 1 2; do for c in 65536 131072 262144; do
  REYNA_B200_PIPELINE_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> $O/chunk.log
  done; done
  cat $O/chunk.log
  ;;
mallopt)
  set -x
  O=gpurun_out/r02g; mkdir -p $O
  for rep in 1 2; do for m in 0 1; do
  REYNA_B200_MALLOPT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); print('mallopt $m', 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> $O/mallopt.log
  done; done
  cat $O/mallopt.log
  ;;
tail)
  set -x
  O=gpurun_out/r02g; mkdir -p $O
  timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
  import json,sys; d=json.loads(sys.stdin.read()); print('e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> $O/tail.log
  nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv >> $O/tail.log
  cat $O/tail.log
  ;;
*) echo "usage: $0 {ab|solver_ab|solver_ab2|window|e2e|arena|minb|setup_diag|chunk|mallopt|tail}";;
esac
