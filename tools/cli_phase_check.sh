# drop-in end to end on the bench workload: phase timings, identity of the .txt with the reference's
set -e
mkdir -p /tmp/m
./tools/gt16_synth --index /tmp/m/h --snapshot /tmp/m/t --contigs chr20:64444167 --seed 1 2>/dev/null
for i in 1 2 3; do ./16gt_b200/bin/snapshotSnpcaller -i /tmp/m/h -o /tmp/m/t -v | grep -E "Read snapshot|Elapsed|Total Running|Called" | tr '\n' ' '; echo; done
sha256sum /tmp/m/t.txt
