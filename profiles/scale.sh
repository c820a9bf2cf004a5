# usage: bash profiles/scale.sh "1 2 4"   (run inside gpurun --gpus N with N >= the largest entry)
for n in ${1:-1 2 4 8}; do
  if [ $n = 1 ]; then timeout 200 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err;
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  python -c "
import json,sys
l=[x for x in open('gpurun_out/scale_$n.json') if x.startswith('{')]
d=json.loads(l[-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('cpu_affinity'), d['clocks'])"
done
