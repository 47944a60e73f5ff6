set -x
O=gpurun_out/r02g; mkdir -p $O
for rep in 1 2; do for m in 0 1; do
REYNA_B200_MALLOPT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mallopt $m', 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> $O/mallopt.log
done; done
cat $O/mallopt.log
