#!/bin/bash
# A/B: HEAD build (ab_head/) vs working tree, same box, alternating
for rep in 1 2; do
for D in ab_head .; do
  for DT in f64 f32; do
  (cd $D && HB_PRE=${HB_PRE_AB:-0} timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --dtype $DT 2>/dev/null \
      | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$D $DT', 'ms/step', round(l['ms_per_step'],4), 'eager', round(l['ms_per_step_eager_launches'],4), 'sections', {k: round(v,4) for k,v in l['section_ms_per_step'].items()})")
  done
done
done
