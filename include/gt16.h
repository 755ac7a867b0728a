/* gt16.h -- C ABI of libgt16.so: the B200 (sm_100a) implementation of 16GT's calling hot
 * path (per-site 16-genotype likelihood, posterior, candidate selection, strand / base-quality
 * bias, window features).  Plain pointers and sizes only; every entry point returns 0 on
 * success or a negative GT16_E* code, and gt16_last_error() gives the text.
 *
 * There is NO CPU implementation behind these calls: gt16_ctx_create() fails with
 * GT16_ENODEVICE when no CUDA device is usable, and every compute entry needs a context.
 *
 * Two levels, mirroring SURVEY.md 8(b):
 *   (1) operator seam  -- the five functions of the reference's vestigial GPU seam
 *       (FisherExactTest.h:10-23), so the reference harness can be A/B-tested by swapping
 *       FisherExactTest.cpp for a 40-line shim;
 *   (2) fused path     -- raw snapshot blocks in, called sites out; this is what the
 *       snapshotSnpcaller drop-in binary and bench.py use.
 */
#ifndef GT16_H_
#define GT16_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT16_OK 0
#define GT16_EINVAL (-1)     /* bad argument */
#define GT16_ENODEVICE (-2)  /* no usable CUDA device / wrong architecture */
#define GT16_ECUDA (-3)      /* CUDA runtime error (text in gt16_last_error) */
#define GT16_EIO (-4)        /* file could not be read / written / parsed */
#define GT16_EDATA (-5)      /* snapshot is internally inconsistent */
#define GT16_ESTATE (-6)     /* call order violated (e.g. compute before prefill) */
#define GT16_EGENOTYPE (-7)  /* "Error in selecting Genotype" (SNP_Caller.cpp:235-239) */

#define GT16_BATCH_SITES 8388608u   /* VariantCaller.cpp:526-528 */

typedef struct gt16_ctx gt16_ctx;

/* ---- context ------------------------------------------------------------------------ */
int gt16_device_count(int *n);                      /* usable sm_100 devices */
int gt16_ctx_create(int device, gt16_ctx **out);
void gt16_ctx_destroy(gt16_ctx *ctx);
const char *gt16_last_error(const gt16_ctx *ctx);   /* ctx may be NULL: last create error */

/* ---- (1) operator seam --------------------------------------------------------------- */

/* copyGenotypeTableToGPU(unsigned int *genotype2int)            FisherExactTest.h:10, .cpp:19-21 */
int gt16_copy_genotype_table(gt16_ctx *ctx, const unsigned int genotype2int[16]);
/* copySnpPriorToGPU(double *snpPrior)                           FisherExactTest.h:12, .cpp:23-25 */
int gt16_copy_snp_prior(gt16_ctx *ctx, const double snp_prior[256]);
/* createSnpPriorArray()                                         SNPFunctions.cpp:13-70 */
int gt16_create_snp_prior(double out256[256]);

/* prefill_likelihood_cache_with_p_err(LikelihoodCache *lc, float b_err, float ub_err)
 *                                                               FisherExactTest.h:23, .cpp:61-92
 * Builds the ten 2000x2000 error tables and the 15000-entry chi-square table in HBM.  If
 * lc_host is non-NULL it must point to sizeof(LikelihoodCache) = 10*2000*2000*8 + 15000*8
 * bytes (likelihood_cache.h:13-31) and receives a copy in the reference's layout. */
int gt16_prefill_likelihood_cache(gt16_ctx *ctx, float b_err, float ub_err, void *lc_host);

/* Start-up of the fused path = what main + selectPossibleSNPs do before the position loop
 * (snapshotSnpcaller.cpp:229-231, VariantCaller.cpp:506-513): prefill with the ini's
 * BalanceSubError / UnbalanceSubError (passed through float like the reference does), default
 * genotype2int + prior, and -- when use_decision_tables != 0 -- the exact memoised decision
 * bitmaps the streaming kernel consults (built on the device by the same evaluator). */
int gt16_configure(gt16_ctx *ctx, double balance_sub_error, double unbalance_sub_error, int use_decision_tables);
/* device milliseconds spent by the last prefill / decision-table build */
int gt16_setup_times(const gt16_ctx *ctx, float *prefill_ms, float *decision_ms);

/* computeGenotypeLikelihoodWrapper(InputStat *in, MetaReference *reference, GenotypeLikelihood *lh,
 *                                  IniParams ini, unsigned n, LikelihoodCache *lc)
 *                                                               FisherExactTest.h:14, .cpp:291-301
 * w points at in[0].W (6 x u32 per site, stride in_stride bytes = sizeof(InputStat) = 144);
 * ref_char points at reference[0].refChar (stride sizeof(MetaReference) = 270);
 * lh receives n x 16 doubles (GenotypeLikelihood, 128 B).  balance/unbalance are
 * ini.Ini_BalanceSubError / Ini_UnbalanceSubError. */
int gt16_compute_genotype_likelihood(gt16_ctx *ctx, const void *w, size_t in_stride, const void *ref_char,
                                     size_t ref_stride, double *lh, uint32_t n, double balance_sub_error,
                                     double unbalance_sub_error);

/* computeGenotypeWrapper(unsigned char *refChar, GenotypeLikelihood *lh, SnpCallingInfo *info,
 *                        IniParams ini, unsigned n)            FisherExactTest.h:17, .cpp:511-521
 * info points at n SnpCallingInfo records of info_stride bytes (168): pD @0, genotype @8,
 * flh[16] @16, bestD @144, secondD @152 are written; the other fields are left untouched. */
int gt16_compute_genotype(gt16_ctx *ctx, const unsigned char *ref_char, const double *lh, void *info,
                          size_t info_stride, uint32_t n);

/* computeStrandBias / computeBaseQualityBias (SNPFunctions.cpp:145-159) for n sites:
 * w,f,r point at 4 x u32 per site with the given stride; sb and bqb receive n x 4 floats. */
int gt16_compute_bias(gt16_ctx *ctx, const void *w, const void *f, const void *r, size_t stride, uint32_t n,
                      float *sb, float *bqb);

/* ---- (2) fused path ------------------------------------------------------------------- */

/* Host-side view of one snapshot (what readSnpInfoSnapshot fills, SnapshotHandler.cpp:259-347)
 * plus the packed reference (HSP.packedDNA). */
typedef struct gt16_snapshot_view {
  const uint64_t *counters;        /* SnpCounter[text_length], 8 bytes each (coreStruct.h:10-30) */
  uint32_t text_length;            /* dnaLength */
  const uint32_t *invalid_bitmap;  /* (text_length+31)>>5 words, bit 31-(p&31); NULL = none */
  uint32_t n_overflow_regions;     /* numOfCPUThreads stored in the snapshot */
  const void *const *overflow;     /* per region: SnpOverflowCounter[] (32 B each) */
  const uint32_t *overflow_size;   /* per region: number of records */
  const uint8_t *pool;             /* memory pool bytes */
  uint32_t pool_bytes;             /* curPtr: bytes in use */
  const uint32_t *packed_ref;      /* 16 bases / word, first base in bits 31..30 */
  uint32_t packed_ref_words;
  /* A sparse (mode 2) snapshot may be passed as it lies in the file instead of expanded: counters == NULL and
   * sparse_entries -> n_sparse_entries packed records {u32 position; SnpCounter} of 12 bytes, positions strictly
   * ascending (bam2snapshot.cpp:527-589 writes them so); a position without a record holds a zero counter.  The records
   * are copied to HBM as they are (12 B per covered position instead of 8 B per genome position) and expanded there. */
  const void *sparse_entries;
  uint64_t n_sparse_entries;
} gt16_snapshot_view;

typedef struct gt16_region { uint32_t start, end; } gt16_region;   /* [start,end), struct.h:5-13 */

typedef struct gt16_indel {
  uint8_t type;      /* 'I', 'D' or '*' */
  uint8_t len;
  uint16_t hq, lq;
  uint8_t seq[64];   /* 2-bit bases, interpreter.cpp:85-87 order */
} gt16_indel;

/* One called site = everything one row of <prefix>.txt needs (interpreter.cpp:107-156),
 * i.e. the payload of the reference's packed .tmpresult record (SNP_Caller.cpp:101-209). */
typedef struct gt16_called_site {
  uint32_t amb;            /* concatenated-coordinate position */
  uint8_t ref;             /* 0..3 */
  uint8_t is_indel;        /* 0, 2 or 3 (isValid of an indel genotype row) */
  uint8_t genotype;        /* 0..15 */
  uint8_t rq_assigned;     /* 0: reference prints the previous row's LQ/RQ (SNP_Caller.cpp:338-353) */
  uint32_t W[6], F[6], R[6];
  double lh[16];           /* GenotypeLikelihood */
  double flh[16];          /* SnpCallingInfo.flh */
  double pD, bestD, secondD;
  float sb[3][4], bqb[3][4];
  float rpb, lgc, rgc, asc, lq, rq;
  uint8_t lg, rg;
  uint16_t polyrun;
  uint32_t total_hq, total_lq;
  gt16_indel indel[2];
} gt16_called_site;

typedef struct gt16_dev_snapshot gt16_dev_snapshot;
typedef struct gt16_calls gt16_calls;

typedef struct gt16_stats {
  uint64_t sites;              /* processed positions */
  uint64_t flagged;            /* sites the streaming kernel sent to exact evaluation */
  uint64_t called;             /* Called Entries */
  uint32_t kernel_launches;    /* kernels of this library launched by the call */
  float ms_h2d, ms_stream, ms_scan, ms_eval, ms_candidates, ms_d2h, ms_total_device;
} gt16_stats;

/* Copy a snapshot to HBM (pinned staging + cudaMemcpyAsync on the copy stream). */
int gt16_snapshot_upload(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot **out);
void gt16_snapshot_release(gt16_ctx *ctx, gt16_dev_snapshot *snap);
/* Copy to HBM what calls on the processed-order range [site_begin, site_end) of `regions` need: that range's counters
 * (8 B/site) and the side arrays its positions can reach -- one shard of a region-sharded genome.  Calls on the
 * returned snapshot must stay inside that range (GT16_EINVAL otherwise). */
int gt16_snapshot_upload_range(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions, uint32_t n_regions,
                               uint64_t site_begin, uint64_t site_end, gt16_dev_snapshot **out);

/* Call every site of the processed-order range [site_begin, site_end) of the concatenated
 * region list (regions == NULL: the single region [0,text_length), snapshotSnpcaller.cpp
 * default).  site_begin must be a multiple of GT16_BATCH_SITES so that the reference's
 * per-batch window semantics are reproduced.  Requires a prior
 * gt16_prefill_likelihood_cache().  Results are ordered by processed index. */
int gt16_call_sites(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                    uint64_t site_begin, uint64_t site_end, gt16_calls **out, gt16_stats *stats);

/* Same call with the result left in HBM (inputs resident, outputs resident): for callers that consume the records
 * on the device or fetch them later.  The handle refers to storage owned by the context and is valid until the
 * next call on that context; gt16_dev_calls_fetch() copies the records to the host. */
typedef struct gt16_dev_calls gt16_dev_calls;
int gt16_call_sites_device(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                           uint64_t site_begin, uint64_t site_end, gt16_dev_calls **out, gt16_stats *stats);
/* The same call, returning as soon as the work has been ENQUEUED: the whole call -- streaming pass, exact evaluation,
 * ordered placement, records, Fisher tests -- is stream-ordered on the context without one host round trip (record
 * counts live on the device; the kernels behind the evaluator read them there), so the host prepares the next call
 * while the GPU runs this one.  The records of a call stay valid until the next call on the context, as with
 * gt16_call_sites_device.  stats at return: sites and kernel_launches; called, flagged, the ms_* fields and any error
 * found on the device arrive with gt16_dev_calls_wait (implied by gt16_dev_calls_count / _device_ptr / _fetch), which
 * may be called until 128 further calls have been made on the context.  Record buffers are sized by a bound (one
 * record per 128 sites + 65,536); a call that finds more is repeated, blocking, inside gt16_dev_calls_wait with
 * buffers of the size it needs. */
int gt16_call_sites_device_nowait(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                                  uint64_t site_begin, uint64_t site_end, gt16_dev_calls **out, gt16_stats *stats);
int gt16_dev_calls_wait(gt16_ctx *ctx, gt16_dev_calls *calls, gt16_stats *stats);
uint64_t gt16_dev_calls_count(const gt16_dev_calls *calls);            /* completes a pending call first */
const void *gt16_dev_calls_device_ptr(const gt16_dev_calls *calls);   /* gt16_called_site[count] in device memory, ordered by site */
int gt16_dev_calls_fetch(gt16_ctx *ctx, const gt16_dev_calls *calls, gt16_calls **out);
void gt16_dev_calls_free(gt16_dev_calls *calls);

/* Same, from HOST buffers: counter blocks are streamed into HBM on the copy stream while the
 * previous block is being evaluated; includes the device->host copy of the calls. */
int gt16_call_sites_host(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions,
                         uint32_t n_regions, uint64_t site_begin, uint64_t site_end, gt16_calls **out,
                         gt16_stats *stats);

/* The same with the records written straight into the caller's buffer `dst` (room for dst_capacity records; page-locked
 * memory -- cudaHostAlloc / cudaHostRegister, e.g. a shared-memory slab that another process reads -- makes the copies
 * asynchronous DMA).  *n_out = records written; GT16_EINVAL with *n_out = records found when dst is too small. */
int gt16_call_sites_host_to(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions, uint32_t n_regions,
                            uint64_t site_begin, uint64_t site_end, gt16_called_site *dst, uint64_t dst_capacity,
                            uint64_t *n_out, gt16_stats *stats);

uint64_t gt16_calls_count(const gt16_calls *calls);
const gt16_called_site *gt16_calls_data(const gt16_calls *calls);
void gt16_calls_free(gt16_calls *calls);

/* ---- host-side formats (the callers / data formats on either side of the path) ---------- */
typedef struct gt16_host_snapshot gt16_host_snapshot;
/* readSnpInfoSnapshot (SnapshotHandler.cpp:259-347); buffers are page-locked when pinned != 0 */
int gt16_snapshot_read(const char *path, int pinned, gt16_host_snapshot **out, char *err, size_t err_len);
/* The same with the counters of positions [pos_lo, pos_hi) only (everything else -- overflow arrays, bitmap, pool -- is
 * read whole): what one shard of a region-sharded genome keeps in (page-locked) memory.  The view's counters pointer
 * stays indexed by genome position; only the window may be dereferenced, which is all that gt16_call_sites_host /
 * gt16_snapshot_upload_range touch for a site range whose positions lie at least 4,608 positions inside it (or at the
 * genome's ends). */
int gt16_snapshot_read_range(const char *path, int pinned, uint32_t pos_lo, uint32_t pos_hi, gt16_host_snapshot **out,
                             char *err, size_t err_len);
/* The same with flags: GT16_READ_PINNED page-locks the buffers; GT16_READ_KEEP_SPARSE leaves the counters of a sparse
 * (mode 2) snapshot as the packed records they are in the file (view: counters == NULL, sparse_entries set) instead of
 * expanding them into a position-indexed array -- gt16_call_sites_host / gt16_snapshot_upload* expand them in HBM.  A
 * dense snapshot is read as before whatever the flag. */
#define GT16_READ_PINNED 1
#define GT16_READ_KEEP_SPARSE 2
int gt16_snapshot_read_ex(const char *path, int flags, uint32_t pos_lo, uint32_t pos_hi, gt16_host_snapshot **out,
                          char *err, size_t err_len);
void gt16_snapshot_view_of(const gt16_host_snapshot *hs, gt16_snapshot_view *view);  /* packed_ref left NULL */
void gt16_snapshot_free(gt16_host_snapshot *hs);

typedef struct gt16_index gt16_index;
/* <prefix>.pac/.ann/.tra as HSPLoad reads them (readIndex.cpp:24-60,96-115) */
int gt16_index_read(const char *prefix, gt16_index **out, char *err, size_t err_len);
uint32_t gt16_index_length(const gt16_index *ix);
uint32_t gt16_index_sequences(const gt16_index *ix);   /* number of sequences (numOfSeq) */
const uint32_t *gt16_index_packed(const gt16_index *ix, uint32_t *n_words);
/* Page-lock the packed reference in place (optional; worthwhile when the index serves many gt16_call_sites_host
 * calls: the copy to HBM then is an asynchronous DMA).  GT16_ECUDA when there is no device. */
int gt16_index_pin(gt16_index *ix);
/* getChrAndPos (HSP.h:273-275): concatenated position -> (sequence name, 1-based position) */
const char *gt16_index_translate(const gt16_index *ix, uint32_t amb, uint32_t *chr_pos);
/* getChrIDFromName + getAmbPos (indexFunction.cpp:6-66): sequence name (with the reference's "chr" prefix
 * leniency) and 1-based position -> concatenated position; returns 0 and *amb, or -1 when there is no such
 * position (unknown sequence, beyond its end, inside a removed stretch) */
int gt16_index_locate(const gt16_index *ix, const char *chr_name, uint32_t chr_pos, uint32_t *amb);
/* its two halves: getChrIDFromName (1-based id, -1 when unknown) and getAmbPos */
int gt16_index_sequence_id(const gt16_index *ix, const char *chr_name);
int gt16_index_locate_id(const gt16_index *ix, int chr_id, uint32_t chr_pos, uint32_t *amb);
void gt16_index_free(gt16_index *ix);
/* Test hook: the emitter's own "%.4e" (kind 0) / "%.4f" (kind 1) formatting of v into out (>= 64 bytes); returns the
 * length.  Must equal printf's output byte for byte. */
int gt16_format_number(double v, int kind, char *out);

/* printHeader + printPossibleSNP (interpreter.cpp:8-31,107-156): byte-identical layout */
int gt16_write_txt(const char *path, const gt16_index *ix, const gt16_called_site *sites, uint64_t n, int append,
                   float *lq_rq_carry /* [2] in/out, may be NULL */);

/* ---- the producer in front of the path: aligned reads -> snapshot (bam2snapshot's counter construction) ---------- */

/* Aligned reads as bamToAlignment (bam2snapshot.cpp:79-143) hands them to updateSnpCounterForReads (SNP.cpp:438), as
 * host arrays.  Reads that the reference skips (unmapped, duplicates, MAPQ below the threshold, length mismatch with
 * the CIGAR; bam2snapshot.cpp:84-186) are not in the list.  Any order; the order decides which of two equally deep indel
 * patterns of a position is listed first, as it does in the reference. */
typedef struct gt16_reads {
  uint64_t n_reads;
  const uint32_t *start;        /* concatenated-coordinate position of the first aligned base (getAmbPos of POS) */
  const uint8_t *reverse;       /* 1: the read is on the reverse strand (FLAG & 16) */
  const uint32_t *length;       /* l_qseq */
  const uint64_t *seq_off;      /* first base of the read in bases[] / quals[] */
  const uint64_t *cigar_off;    /* first operation of the read in cigar[] */
  const uint32_t *n_cigar;
  uint64_t n_bases, n_cigar_ops;
  const uint8_t *bases;         /* 0..3 = A C G T as in SEQ, i.e. in reference orientation (N and IUPAC codes as bamcode2snpcode maps them, bam2snapshot.h:52-58) */
  const uint8_t *quals;         /* phred scores as in QUAL (0xFF: absent) */
  const uint32_t *cigar;        /* BAM encoding: length << 4 | op, op = 0 M, 1 I, 2 D, 3 N, 4 S, 5 H, 6 P, 7 =, 8 X */
} gt16_reads;

/* what loadBam2snapshotParas (bam2snapshot.cpp:336-369) derives from bam2snapshot.ini */
typedef struct gt16_pileup_params {
  uint8_t weight_map[41];          /* quality -> weight 0..4 (Weight0..Weight3) */
  uint32_t trim_size;              /* statTrimSize: bases ignored at either end of a read */
  uint32_t indel_weight_threshold; /* from IndelQualityThreshold */
  uint32_t n_overflow_regions;     /* NumOfCpuThreads: number of overflow arrays of the snapshot */
} gt16_pileup_params;
/* defaults of the shipped bam2snapshot.ini (Weight 10/13/17/20, trim 5, IndelQualityThreshold 13, 16 regions) */
void gt16_pileup_default_params(gt16_pileup_params *p);

/* A snapshot built in memory: `view` is what gt16_call_sites_host / gt16_snapshot_upload / gt16_snapshot_write take
 * (packed_ref left NULL).  Released with gt16_built_snapshot_free. */
typedef struct gt16_built_snapshot {
  gt16_snapshot_view view;
  void *overflow_storage;
  /* how the time went: host (sorting / packing the reads), host->device copies, the k_pileup launch (CUDA events),
   * device->host copies of the result; aligned bases in the reads */
  float ms_pack, ms_h2d, ms_kernel, ms_d2h;
  uint64_t aligned_bases;
  uint32_t heavy_tiles;     /* tiles of 2,048 positions with more than 2,048 indel events: done in a second launch */
} gt16_built_snapshot;

/* updateSnpCounterForReads over all reads + readFilesForOverflow (bam2snapshot.cpp:221-330; SNP.cpp:438-891,893-1688):
 * per-position base counters, overflow promotion, invalid bits, pooled insertion / deletion patterns, built on the GPU
 * as one histogram pass (gt16_pileup.cu).  The decoded content equals the reference producer's; slot order, array
 * indices and pool offsets -- which depend on the order in which the reference's threads get there -- need not.
 * The read records are checked first (start inside the genome, CIGAR and sequence ranges inside the arrays, operations
 * 0..8, no CIGAR longer than its read): GT16_EINVAL names the first read that fails. */
int gt16_build_snapshot(gt16_ctx *ctx, const gt16_reads *reads, uint32_t text_length, const gt16_pileup_params *params,
                        gt16_built_snapshot *out);
void gt16_built_snapshot_free(gt16_built_snapshot *s);

/* writeSnpInfoSnapshot (SnapshotHandler.cpp:83-236): the container with its page-aligned subfiles; dense (mode 1) when
 * at least 10 % of the positions carry data, else sparse (mode 2: positions within 100 of a populated one,
 * bam2snapshot.cpp:527-589).  The pool is written at its used size rounded up to a page, not at 1 GiB. */
int gt16_snapshot_write(const char *path, const gt16_snapshot_view *view, char *err, size_t err_len);

/* SAM text -> gt16_reads with the reference's filters (bam2snapshot.cpp:79-186: MAPQ threshold, unmapped, FLAG 0x400,
 * l_qseq against the CIGAR); sequence names are resolved against the index like getChrIDFromName / getAmbPos. */
typedef struct gt16_sam gt16_sam;
int gt16_sam_read(const char *path, const gt16_index *ix, uint32_t mapq_threshold, gt16_sam **out, char *err, size_t err_len);
/* The same from a BAM file (BGZF inflated with zlib; the records bam_read1 hands to bamToAlignment). */
int gt16_bam_read(const char *path, const gt16_index *ix, uint32_t mapq_threshold, gt16_sam **out, char *err, size_t err_len);
void gt16_sam_reads(const gt16_sam *sam, gt16_reads *reads);
uint64_t gt16_sam_total_alignments(const gt16_sam *sam);   /* lines read, including the skipped ones */
void gt16_sam_free(gt16_sam *sam);

#ifdef __cplusplus
}
#endif
#endif /* GT16_H_ */
