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
