#!/bin/bash
# profile_round.sh TAG: the ncu evidence of one round, taken on the GPU box (each command has run once without ncu
# before: bench.py lines of the same code are in gpurun_out/).  Writes gpurun_out/TAG_*.
tag=${1:-rX}
cd "$(dirname "$0")/.."
B="python bench.py --workload chr20 --steps 2 --warmup 1 --no-cpu-baseline --no-file"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stream_tma|k_eval_mask|k_rank|k_candidates|k_bias|k_mask_count|k_scan" -s 12 -c 8 -o gpurun_out/${tag}_kernels $B > gpurun_out/${tag}_ncu_kernels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stream_runs|k_stream_regions" -s 16 -c 1 -o gpurun_out/${tag}_exome python bench.py --workload exome --steps 2 --warmup 1 --no-cpu-baseline --no-file > gpurun_out/${tag}_ncu_exome.log 2>&1
# the producer: reads of the 5 Mb / 30x producer config through 16gt_b200/bin/bam2snapshot
python - <<'PY'
import os, subprocess, sys, tempfile
sys.path.insert(0, "tests")
import gt16_testlib as T
wd = tempfile.mkdtemp(prefix="gt16_prof_")
pre = os.path.join(wd, "p")
subprocess.run([T.READSIM, "--out", pre] + T.PRODUCER_CONFIGS["prod_5mb"]["readsim"], check=True, stderr=subprocess.DEVNULL)
subprocess.run([T.REF_MKINDEX, pre + ".fa", pre + ".index"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
open("gpurun_out/.prof_pileup_cmd", "w").write("%s -i %s -s %s -o %s" % (os.path.join(T.PKG, "bin", "bam2snapshot"), pre + ".index", pre + ".sam", pre + ".mine"))
PY
$(cat gpurun_out/.prof_pileup_cmd) > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pileup" -c 1 -o gpurun_out/${tag}_pileup $(cat gpurun_out/.prof_pileup_cmd) > gpurun_out/${tag}_ncu_pileup.log 2>&1
rm -f gpurun_out/.prof_pileup_cmd
