# 1/2/4/8-GPU weak-scaling run of bench.py on one box (the driver's launch line), results under gpurun_out/
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_v4_n$n.json 2> gpurun_out/scale_v4_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_v4_n$n.json 2> gpurun_out/scale_v4_n$n.err; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_v4_n$n.json')); print($n, round(d['ms_per_step'],3), '%.3e' % d['value'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and '%.3e' % d['e2e']['value'])"
done
