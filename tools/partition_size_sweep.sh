for E in 1048576 524288 262144 131072; do
  timeout 300 python bench.py --elements $E --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($E, 'step_ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'], 'e2e_s', round(d['e2e']['s_per_step'],4))"
done
