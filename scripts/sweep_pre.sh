#!/bin/bash
# A/B of the k_step "pre" roles at several per-GPU task counts (one GPU)
DT=${DT:-f64}
for T in ${TS:-32 64 128 256}; do
  for P in 0 1; do
    HB_PRE=$P timeout 120 python bench.py --tasks $T --steps 50 --warmup 5 --no-cpu-baseline --dtype $DT 2>/dev/null \
      | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$DT T=$T pre=$P', 'ms/step', round(l['ms_per_step'],4), 'eager', round(l['ms_per_step_eager_launches'],4), 'sections', {k: round(v,4) for k,v in l['section_ms_per_step'].items()}, 'loss', l['final_loss'])"
  done
done
