#!/bin/bash
# profile_round.sh TAG: the ncu evidence of one round, taken on the GPU box (each command has run once without ncu
# before: bench.py lines of the same code are in gpurun_out/).  Writes gpurun_out/TAG_*.
tag=${1:-rX}
cd "$(dirname "$0")/.."
B="python bench.py --workload chr20 --steps 2 --warmup 1 --no-cpu-baseline --no-file"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stream_tma|k_eval_mask|k_rank|k_candidates|k_bias|k_mask_count|k_scan" -s 12 -c 8 -o gpurun_out/${tag}_kernels $B > gpurun_out/${tag}_ncu_kernels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stream_regions|k_scatter_sparse|k_zero_spans" -s 2 -c 3 -o gpurun_out/${tag}_exome python bench.py --workload exome --steps 2 --warmup 1 --no-cpu-baseline --no-file > gpurun_out/${tag}_ncu_exome.log 2>&1
