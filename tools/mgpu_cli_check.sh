set -e
mkdir -p /tmp/m
./tools/gt16_synth --index /tmp/m/g --snapshot /tmp/m/s --genome-len 8600000 --n-contigs 2 --snp-rate 4e-5 --indel-rate 1e-5 --seed 17 2>/dev/null
./16gt_b200/bin/snapshotSnpcaller -i /tmp/m/g -o /tmp/m/s -v | grep -E "GT16|Called|Total Running"
python - <<PY
import sys
sys.path.insert(0,"tests")
import gt16_testlib as T
got=open("/tmp/m/s.txt","rb").read()
print("2-GPU CLI two_batch vs golden: identical bytes:", got==T.golden_txt("two_batch"))
PY
./tools/gt16_synth --index /tmp/m/h --snapshot /tmp/m/t --contigs chr20:64444167 --seed 1 2>/dev/null
./16gt_b200/bin/snapshotSnpcaller -i /tmp/m/h -o /tmp/m/t -v | grep -E "GT16|Called|Total Running|Elapsed|Read snapshot"
sha256sum /tmp/m/t.txt
