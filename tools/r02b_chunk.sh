set -x
O=gpurun_out/r02g; mkdir -p $O
for rep in 1 2; do for c in 65536 131072 262144; do
REYNA_B200_PIPELINE_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', 'e2e_s', round(d['e2e']['s_per_step'],4), d['e2e']['breakdown_last_step'])" >> $O/chunk.log
done; done
cat $O/chunk.log
