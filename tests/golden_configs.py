"""Small parity configs shared by tests/golden/make_golden.py (which runs the UNMODIFIED reference
objects in oracle/_ref on them, here, where /root/reference exists) and by the tests (which
regenerate the same inputs with tools/gt16_synth -- a pure function of its arguments -- and
compare against the committed reference outputs in tests/golden/).

Each entry mirrors one of BASELINE.json's configs at a size the reference finishes in seconds:
  wgs_small    config 1/2: 30x WGS, planted SNPs/indels, two contigs
  exome_small  config 4  : -e region.bin, sparse (mode 2) snapshot, 150x targets
  stress_small config 5  : 60x, dense indels with multi-pattern (pooled) lists, long indels,
                           depth hotspots (overflow promotion, W >= 2000 -> chi-square fallback,
                           invalid bits), odd number of overflow regions, coverage gaps
  two_batch    config 3  : > 8,388,608 sites, so the per-batch window reset and the batch-aligned
                           sharding of SURVEY 8(e) are exercised
"""

CONFIGS = {
    "wgs_small": dict(
        synth=["--contigs", "chrA:180000,chrB:70000", "--seed", "7"], exome=False),
    "exome_small": dict(
        synth=["--genome-len", "3000000", "--n-contigs", "3", "--exome", "300:300", "--depth", "150",
               "--seed", "11"], exome=True),
    "stress_small": dict(
        synth=["--genome-len", "400000", "--n-contigs", "2", "--depth", "60", "--indel-rate", "2e-3",
               "--multi-pattern-frac", "0.5", "--long-indel-frac", "0.1", "--hotspot-rate", "2e-4",
               "--hotspot-depth", "1500", "--invalid-rate", "1e-3", "--ovf-regions", "7", "--gaps", "2:5000",
               "--seed", "13"], exome=False),
    # same snapshot as stress_small, called with a non-default ini (ten distinct error tables, other decision
    # bitmaps): pins the oracle and the GPU path against the reference away from the shipped error rates
    "stress_small_ini2": dict(
        synth=["--genome-len", "400000", "--n-contigs", "2", "--depth", "60", "--indel-rate", "2e-3",
               "--multi-pattern-frac", "0.5", "--long-indel-frac", "0.1", "--hotspot-rate", "2e-4",
               "--hotspot-depth", "1500", "--invalid-rate", "1e-3", "--ovf-regions", "7", "--gaps", "2:5000",
               "--seed", "13"], exome=False, ini=dict(BalanceSubError=0.02, UnbalanceSubError=0.003)),
    "two_batch": dict(
        synth=["--genome-len", "8600000", "--n-contigs", "2", "--snp-rate", "4e-5", "--indel-rate", "1e-5",
               "--seed", "17"], exome=False),
}

# BASELINE.json configs[1] at FULL size (what bench.py times).  The reference's .txt is 47 MB, so only
# a digest of its call-set columns is committed (tests/golden/MANIFEST.json: callset_sha256).
FULL_CONFIGS = {
    "chr20_full": dict(synth=["--contigs", "chr20:64444167", "--seed", "1"], exome=False),
    # config 5 flavour at chr20 size: 60x, indels 1/500 with pooled multi-pattern lists, depth hotspots
    # (overflow promotion, chi-square fallback), invalid bits, coverage gaps; 8 batches
    "stress_full": dict(synth=["--contigs", "chrS:64444167", "--depth", "60", "--indel-rate", "2e-3",
                               "--multi-pattern-frac", "0.4", "--long-indel-frac", "0.05", "--hotspot-rate", "2e-5",
                               "--hotspot-depth", "1500", "--invalid-rate", "1e-4", "--gaps", "6:100000",
                               "--seed", "5"], exome=False),
    # config 4 flavour: 20,000 targets of ~300 bp at 150x on a 256 Mb genome, sparse (mode 2) snapshot, -e region.bin
    "exome_mid": dict(synth=["--genome-len", "256000000", "--n-contigs", "8", "--exome", "20000:300", "--depth", "150",
                             "--seed", "4"], exome=True),
    # config 4 at FULL size: 200,000 targets of ~300 bp at 150x on a 3.1 Gb genome of 24 contigs (positions beyond
    # 2^31), sparse (mode 2) snapshot, -e region.bin
    "exome_full": dict(synth=["--genome-len", "3100000000", "--n-contigs", "24", "--exome", "200000:300", "--depth", "150",
                              "--seed", "6"], exome=True),
}
# configs[2] at FULL size: 3.1 Gb, 24 contigs, 30x, dense snapshot (24.8 GB of counters, 370 batches), coverage gaps.
# Needs ~30 GB of RAM and disk to regenerate: run by tools/full_genome_check.py / GT16_RUN_HUGE=1, not by default.
HUGE_CONFIGS = {
    "wgs_full": dict(synth=["--genome-len", "3100000000", "--n-contigs", "24", "--gaps", "8:100000", "--seed", "3"],
                     exome=False),
    # configs[4] at FULL size: 3.1 Gb at 60x, indels 1/500 with pooled multi-pattern lists, depth hotspots (overflow
    # promotion, chi-square fallback), invalid bits, coverage gaps
    "stress_huge": dict(synth=["--genome-len", "3100000000", "--n-contigs", "24", "--depth", "60", "--indel-rate", "2e-3",
                               "--multi-pattern-frac", "0.4", "--long-indel-frac", "0.05", "--hotspot-rate", "2e-5",
                               "--hotspot-depth", "1500", "--invalid-rate", "1e-4", "--gaps", "12:100000",
                               "--seed", "9"], exome=False),
}
FULL_CONFIGS.update(HUGE_CONFIGS)

# Snapshots written by the reference's OWN producer (tests/golden/make_producer_golden.py): tools/gt16_readsim makes a
# FASTA and aligned reads (SAM), oracle/_ref/ref_mkindex and oracle/_ref/ref_bam2snapshot -- the unmodified reference
# tools -- turn them into index files and a snapshot.
#   prod_5mb       BASELINE.json configs[0]: 5 Mb, three contigs, 30x diploid, planted SNPs (1/1000, het:hom 2:1) and
#                  indels (1/8000, 1-12 bp plus 13-30 bp ones that need pooled patterns, compound hets), soft clips,
#                  duplicates, three 1500x hotspots (overflow promotion, chi-square fallback) and one 20000x hotspot
#                  (u16 counters saturate: invalid bits)
#   prod_two_batch more than 8,388,608 sites (two batches: per-batch window reset, batch-aligned sharding) at 12x
PRODUCER_CONFIGS = {
    "prod_5mb": dict(readsim=["--contigs", "chrA:3000000,chrB:1500000,chrC:500000", "--hotspots", "3:300:1500,1:120:20000",
                              "--seed", "1"]),
    "prod_two_batch": dict(readsim=["--contigs", "chrA:5000000,chrB:3700000", "--depth", "12", "--snp-rate", "1e-4",
                                    "--indel-rate", "2e-5", "--hotspots", "1:200:1200", "--seed", "2"]),
}
CONFIGS_ALL = dict(CONFIGS, **FULL_CONFIGS)


def synth_command(synth_bin, name, workdir):
    """argv that regenerates config `name` under workdir; returns (argv, index_prefix, out_prefix, region_bin)"""
    import os
    cfg = CONFIGS_ALL[name]
    idx = os.path.join(workdir, name + ".idx")
    out = os.path.join(workdir, name)
    region = os.path.join(workdir, name + ".region.bin") if cfg["exome"] else None
    argv = [synth_bin, "--index", idx, "--snapshot", out] + cfg["synth"]
    if region:
        argv += ["--region-out", region]
    return argv, idx, out, region
