# One-off large-scale parity check (config-3 flavour at ~1/3 scale): 1.0 Gb, 8 contigs, coverage gaps, 120 batches,
# > 2^28 sites in one call.  The UNMODIFIED reference objects and the drop-in binary on the same snapshot file;
# the two .txt files must have the same sha256.
set -e
D=${1:-/tmp/big}
mkdir -p $D
df -h $D | tail -1
time ./tools/gt16_synth --index $D/g --snapshot $D/s --genome-len 1000000007 --n-contigs 8 --snp-rate 2e-4 --indel-rate 2e-5 --gaps 5:100000 --seed 21 --threads 16
ls -la $D | head
cp oracle/_ref/ref_caller.ini $D/ 2>/dev/null || true
( time ./16gt_b200/bin/snapshotSnpcaller -i $D/g -o $D/s -v ) 2>&1 | grep -E "GT16|Called|Total Running|Read snapshot|Elapsed|real"
sha256sum $D/s.txt | tee $D/ours.sha
mv $D/s.txt $D/ours.txt
( time ./oracle/_ref/ref_caller $D/g $D/s -t 16 ) 2>&1 | grep -E "REF_|Called|Total Running|real"
sha256sum $D/s.txt | tee $D/ref.sha
cmp $D/s.txt $D/ours.txt && echo "LARGE GENOME: .txt files are byte-identical"
wc -l $D/s.txt
