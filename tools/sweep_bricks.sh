#!/bin/bash
# tuning sweep for k_integrate_bricks (not a bench): prints the 2integrate stage time per configuration
for cfg in "9 6 2 320" "9 6 3 320" "9 6 4 320" "25 6 2 320" "13 6 2 320" "5 6 2 320" "9 12 3 320" "9 4 3 160" "9 8 3 160" "13 8 3 160" "25 8 3 160" "9 8 4 160" "9 8 4 128"; do
  set -- $cfg
  RR_BRICK_ZCHUNK=$1 RR_BRICK_GRID=$2 RR_BRICK_MINB=$3 RR_BRICK_THREADS=$4 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['stages_ms'], d['ms_per_step'])"
done
