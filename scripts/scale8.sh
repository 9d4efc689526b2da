#!/bin/bash
O=gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline 2>> $O/r1p_scale.err | grep '^{' > $O/r1p_bench_n8.json
python -c "import json; l=json.loads(open('$O/r1p_bench_n8.json').read()); print('N=8', round(l['value'],1), 'steps/s', round(l['ms_per_step'],4), 'ms eager', round(l['ms_per_step_eager_launches'],4), 'e2e', round(l['e2e']['value'],1), 'loss', l['final_loss'])"
