#!/bin/bash
# strong-scaling runs on one 8-GPU box (bench.py contract; rank 0 prints the line)
O=gpurun_out
TAG=${TAG:-r1m}
run() { # N dtype
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29520 + $1)) bench.py --gpus $1 --steps 100 --warmup 5 --no-cpu-baseline --dtype $2 2>> $O/${TAG}_scale.err | grep '^{' > $O/${TAG}_bench_n$1_$2.json
  python -c "import json,sys; l=json.loads(open('$O/${TAG}_bench_n$1_$2.json').read()); print('N=$1 $2', round(l['value'],1), 'steps/s', round(l['ms_per_step'],4), 'ms  e2e', round(l['e2e']['value'],1), 'loss', l['final_loss'])"
}
run 8 f64
run 4 f64
run 8 f32
