#!/bin/bash
# sweep_variants.sh NAME ...: device-step breakdown for each tools/probes/variants/libgt16_NAME.so
cd "$(dirname "$0")/../.."
for v in "$@"; do
  GT16_LIB=$PWD/tools/probes/variants/libgt16_$v.so python bench.py --workload ${GT16_SWEEP_WORKLOAD:-chr20} --steps 20 --warmup 3 --no-cpu-baseline --no-file 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); b=d['device_ms_breakdown_rank0']; print('$v', 'stream %.4f eval %.4f scan %.4f cand %.4f' % (b['ms_stream'], b['ms_eval'], b['ms_scan'], b['ms_candidates']), 'step %.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
done
