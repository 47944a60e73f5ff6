set -x
mkdir -p gpurun_out/r02b
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02b/parity_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/parity_tests.log; tail -2 gpurun_out/r02b/parity_tests.log
for i in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run $i', 'step', round(d['ms_per_step'],4), 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> gpurun_out/r02b/e2e.log
done
cat gpurun_out/r02b/e2e.log
