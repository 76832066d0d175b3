/*
 * rsem_b200.h - C ABI of librsem_b200.so: the B200 (sm_100a) implementation of RSEM's
 * expression-estimation hot path (rsem-run-em E/M loop and rsem-run-gibbs sampler).
 *
 * The reference has no in-process plugin/FFI interface for this path: its boundary is the
 * process boundary of `rsem-run-em` / `rsem-run-gibbs` (argv + files).
 * The drop-in executables in rsem_b200/host/ keep that boundary byte for byte and call ONLY the
 * functions below for compute.  Every entry point cites the reference code it replaces
 * (paths relative to /root/reference).
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; rsem_b200_last_error() returns a
 *     thread-local message.  Host callers print it and exit(-1) (reference convention,
 *     my_assert.h:34-41).
 *   - all pointer arguments are HOST pointers owned by the caller unless stated otherwise;
 *     device buffers are owned by the context.
 *   - calls are synchronous with respect to the host unless stated otherwise.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - transcript ids are the reference's: 0 = noise "transcript", 1..M = isoforms.
 *   - `sid` is signed exactly as SingleHit stores it: |sid| = transcript id, sid < 0 = reverse
 *     strand (SingleHit.h:24-26).
 */
#ifndef RSEM_B200_H_
#define RSEM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSEM_B200_VERSION 100 /* 0.1.0 */

/* error codes */
#define RSEM_B200_OK 0
#define RSEM_B200_ERR_CUDA 1     /* CUDA runtime / driver failure, or no device */
#define RSEM_B200_ERR_ARG 2      /* invalid argument / call order */
#define RSEM_B200_ERR_NCCL 3     /* NCCL failure or libnccl not loadable */
#define RSEM_B200_ERR_UNSUPPORTED 4

typedef struct rsem_b200_ctx rsem_b200_ctx;

/* ------------------------------------------------------------------------------------------------
 * Model tables handed to the device (read-only during an E-step).
 * Replaces the per-thread shared `ModelType& model` read by E_STEP (EM.cpp:176-247):
 *   Orientation.h:12-25, LenDist.h:57-77, RSPD.h:63-75, Profile.h:114-124, QProfile.h:111-120,
 *   NoiseProfile.h:104-113, NoiseQProfile.h:115-124 and `mw` (SingleQModel.h:482-544).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rsem_b200_lendist {
    int32_t lb, ub, span;  /* support is (lb, ub], span = ub - lb (LenDist.h:17-32)            */
    const double* pdf;     /* span + 1 entries, pdf[0] = 0                                     */
    const double* cdf;     /* span + 1 entries                                                 */
} rsem_b200_lendist;

typedef struct rsem_b200_model {
    int32_t model_type;   /* 0 SingleModel, 1 SingleQModel, 2 PairedEndModel, 3 PairedEndQModel (EM.cpp:661-667) */
    int32_t M;            /* number of transcripts                                             */
    int32_t seed_len;     /* ModelParams::seedLen                                              */
    int32_t est_rspd;     /* 0/1                                                               */
    int32_t rspd_B;       /* number of RSPD bins                                               */
    int32_t has_mld;      /* single-end: 1 iff --fragment-length-mean given (mld != NULL);
                             paired-end: always 1 (mld = mate length distribution)             */
    int32_t pro_len;      /* no-quality models: Profile::proLen (= mparams.maxL); else 0       */
    int32_t reserved;
    double ori[2];        /* P(forward), P(reverse)                                            */
    rsem_b200_lendist gld;
    rsem_b200_lendist mld;        /* valid iff has_mld                                         */
    const double* rspd_pdf;       /* B + 2 entries (RSPD.h:21-33)                              */
    const double* rspd_cdf;       /* B + 2 entries                                             */
    const double* profile;        /* Q models: [100][5][5]; no-Q: [pro_len][5][5], p[.][ref][read] */
    const double* noise_profile;  /* Q models: [100][5];    no-Q: [5]                          */
    const double* mw;             /* M + 1 mask weights, mw[0] = 1                             */
} rsem_b200_model;

/* Sufficient statistics produced by one model-update E-step (rounds 1-10).
 * Replaces the helper models `mhps[t]` filled by update/updateNoise (e.g. SingleQModel.h:168-221,
 * PairedEndQModel.h:161-188) and summed by Model::collect (EM.cpp:400-404).
 * The caller provides the buffers; entries the model does not estimate are left untouched.   */
typedef struct rsem_b200_model_stats {
    double* profile;        /* same shape as rsem_b200_model.profile, raw (un-normalised) sums  */
    double* noise_profile;  /* same shape as rsem_b200_model.noise_profile, raw sums WITHOUT the N0 counts `c` */
    double* gld_pdf;        /* paired-end only: gld_span + 1 sums indexed by len - gld_lb      */
    int32_t gld_lb, gld_span; /* support the helper histogram uses: the ORIGINAL (minL-1, maxL] (PairedEndQModel.h:72) */
    double* rspd_pdf;       /* est_rspd only: B + 2 sums (RSPD.h:43-59)                        */
} rsem_b200_model_stats;

/* per-round scalars printed on the reference's "ROUND = ..." line (EM.cpp:406-415) */
typedef struct rsem_b200_round_stats {
    double sum;      /* sum of counts incl. N0  */
    double bchange;  /* biggest relative change */
    int64_t totnum;  /* #transcripts with probv >= 1e-7 and relative change >= 1e-3 */
} rsem_b200_round_stats;

/* ------------------------------------------------------------------------------------------------ */
int rsem_b200_version(void);
const char* rsem_b200_last_error(void);
int rsem_b200_device_count(int* n_devices);

/* One context per GPU (one host thread or process per context). */
int rsem_b200_ctx_create(int device, rsem_b200_ctx** out);
int rsem_b200_ctx_destroy(rsem_b200_ctx* ctx);
/* Run all subsequent work on `cuda_stream` (a cudaStream_t owned by the caller, e.g. a
 * torch.cuda.Stream handle) instead of the context's own stream.  NULL restores the default.  */
int rsem_b200_ctx_set_stream(rsem_b200_ctx* ctx, void* cuda_stream);
int rsem_b200_ctx_sync(rsem_b200_ctx* ctx);
/* bytes of device memory currently held by the context */
int rsem_b200_ctx_device_bytes(rsem_b200_ctx* ctx, uint64_t* bytes);

/* ---- multi-GPU: reads are sharded across contexts; the per-transcript count vector (and, in
 * rounds 1-10, the model sufficient statistics) are summed with ncclAllReduce over NVLink.
 * Replaces the serial merge `countvs[0][j] += countvs[i][j]` (EM.cpp:385-389) and
 * Model::collect (EM.cpp:402).  libnccl.so.2 is dlopen()ed on first use.                       */
#define RSEM_B200_UNIQUE_ID_BYTES 128
int rsem_b200_comm_unique_id(void* id_out /* 128 bytes */);
int rsem_b200_comm_init(rsem_b200_ctx* ctx, const void* id /* 128 bytes */, int n_ranks, int rank);

/* Read sharding over GPUs = the reference's read sharding over threads (init<>, EM.cpp:135-157): shard i gets the
 * contiguous reads [bounds[i], bounds[i + 1]) holding about nHits / n_shards hits, every shard at least one read while
 * reads are left.  Pure host arithmetic (no device needed); bounds has n_shards + 1 entries.                      */
int rsem_b200_shard_reads(uint64_t N, const uint64_t* row_ptr /* N + 1 */, int32_t n_shards, uint64_t* bounds);

/* ---- data upload ------------------------------------------------------------------------------
 * Hit buffer = HitContainer<SingleHit|PairedEndHit> (HitContainer.h:12-59) as SoA CSR.
 * row_ptr has N + 1 entries, row_ptr[0] = 0, row_ptr[N] = H.  insertL may be NULL (single-end).
 * pos/insertL may both be NULL when only frozen-conprb rounds / Gibbs are going to run.        */
int rsem_b200_upload_hits(rsem_b200_ctx* ctx, uint64_t N, uint64_t H, int32_t M, const uint64_t* row_ptr,
                          const int32_t* sid, const int32_t* pos, const int32_t* insertL);
/* Directly set hit.conprb / ncpv (used when resuming from an .ofg-like matrix, by tests and the
 * benchmark).  conprb: H doubles, ncpv: N doubles.  The copies run on a second stream; meanwhile the
 * directory of the equivalence-class layout the frozen-conprb rounds use (it depends on row_ptr / sid
 * only) is built on the device, so a following rsem_b200_em_rounds starts without that latency.    */
int rsem_b200_upload_conprb(rsem_b200_ctx* ctx, const double* conprb, const double* ncpv);
int rsem_b200_download_conprb(rsem_b200_ctx* ctx, double* conprb, double* ncpv);
/* Same as upload_hits + upload_conprb but from DEVICE pointers already resident on ctx's GPU
 * (no copy across PCIe; buffers are copied device-to-device into the context, on the context's
 * stream: work that produces them on another stream must have completed before the call).      */
int rsem_b200_adopt_device_matrix(rsem_b200_ctx* ctx, uint64_t N, uint64_t H, int32_t M, const uint64_t* d_row_ptr,
                                  const int32_t* d_sid, const double* d_conprb, const double* d_ncpv);

/* Reads, parsed once and kept resident (the reference re-parses the FASTA/FASTQ text every round
 * 1-11: ReadReader.h:34-38, EM.cpp:195-202).  mate m of read i occupies
 * base[m][off[m][i] .. off[m][i+1]) with codes A0 C1 G2 T3 N4 (utils.h:36-55); qual holds
 * phred+33 - 33 (QProfile.h:42) or is NULL for no-quality read types.  low_quality[i] is the
 * result of calc_lq (SingleReadQ.h:63-95, PairedEndReadQ.h:58-65).                             */
int rsem_b200_upload_reads(rsem_b200_ctx* ctx, int32_t n_mates, const uint64_t* off1, const uint8_t* base1,
                           const uint8_t* qual1, const uint64_t* off2, const uint8_t* base2, const uint8_t* qual2,
                           const uint8_t* low_quality);
/* Reference transcripts (Refs.h / RefSeq.h): forward-strand base codes of transcript t in
 * seq[seq_off[t] .. seq_off[t] + totLen[t]) for t = 1..M (index 0 unused), seed masks as the
 * 32-bit words of RefSeq::fmasks (RefSeq.h:89-97) at mask_words[mask_off[t] ..).               */
int rsem_b200_upload_refs(rsem_b200_ctx* ctx, int32_t M, const uint64_t* seq_off, const uint8_t* seq,
                          const int32_t* full_len, const int32_t* tot_len, const uint64_t* mask_off,
                          const uint32_t* mask_words);
int rsem_b200_set_model(rsem_b200_ctx* ctx, const rsem_b200_model* model);

/* ---- compute ----------------------------------------------------------------------------------- */
/* K1: hit.conprb = model.getConPrb(read, hit), ncpv = model.getNoiseConPrb(read) for every hit
 * (calcConProbs, EM.cpp:249-278; getConPrb e.g. SingleQModel.h:101-151).                       */
int rsem_b200_calc_conprb(rsem_b200_ctx* ctx);

int rsem_b200_set_theta(rsem_b200_ctx* ctx, const double* theta /* M + 1 */);
int rsem_b200_get_theta(rsem_b200_ctx* ctx, double* theta /* M + 1 */);

/* K2 + (allreduce) + K4, repeated: frozen-conprb EM rounds (ROUND >= 12 of EM.cpp:364-416).
 * Runs rounds first_round, first_round + 1, ... without host round trips until either
 * `max_rounds_this_call` rounds ran or the reference's loop condition
 *     ROUND < min_round || (totNum > 0 && ROUND < max_round)           (EM.cpp:416)
 * became false (the stop test runs on the device).  n0 is N0 of the .cnt file (EM.cpp:392).
 * stats_out receives one entry per executed round; *rounds_run their number; *stopped = 1 if
 * the loop condition ended the run.  With a communicator every rank must make the same call.   */
int rsem_b200_em_rounds(rsem_b200_ctx* ctx, int32_t first_round, int32_t max_rounds_this_call, int32_t min_round,
                        int32_t max_round, double n0, rsem_b200_round_stats* stats_out, int32_t* rounds_run,
                        int32_t* stopped);

/* One model-updating round (ROUND <= 10, EM.cpp:368-404): [K1 if conprb is stale] + K2 with
 * posterior write-back + K3 (update/updateNoise) + allreduce + K4.  The caller then rebuilds the
 * master model from `stats` (Model::finish) and calls rsem_b200_set_model again, which marks
 * conprb stale (needCalcConPrb).                                                                */
int rsem_b200_em_model_round(rsem_b200_ctx* ctx, double n0, rsem_b200_model_stats* stats,
                             rsem_b200_round_stats* round_stats);

/* Final pass with calcExpectedWeights = true (EM.cpp:460-478): E-step with the current theta that
 * overwrites hit.conprb / ncpv with the posteriors and returns the expected counts
 * (M + 1, summed over ranks, WITHOUT adding N0).  theta is not changed.  Afterwards conprb / ncpv
 * hold posteriors (read them with download_conprb): further E-steps on this context need a new
 * upload_conprb / calc_conprb first and fail with RSEM_B200_ERR_ARG otherwise.                   */
int rsem_b200_expected_weights(rsem_b200_ctx* ctx, double* counts_out);

/* ---- Gibbs (rsem-run-gibbs, Gibbs.cpp:265-353) ------------------------------------------------
 * The .ofg matrix: rows of (sid, conprb) with the noise entry (sid 0) first when present
 * (Gibbs.cpp:119-131).                                                                          */
typedef struct rsem_b200_gibbs_params {
    int32_t M;
    int32_t burnin, gap;
    int32_t n_chains;             /* = the reference's nThreads after clamping (Gibbs.cpp:478-481)       */
    const int32_t* chain_samples; /* n_chains entries: samples kept by chain t (Gibbs.cpp:211-223)       */
    const uint32_t* chain_seeds;  /* n_chains entries: mt19937 seeds from engineFactory (sampling.h:19-42) */
    double n0;                    /* N0 from the .ofg header                                             */
    const int32_t* init_counts;   /* M + 1: 0, or -1 for omitted transcripts (Gibbs.cpp:152-167)         */
    const double* pseudo_counts;  /* M + 1: pseudoC or the --prior values (Gibbs.cpp:298-306)            */
    double totc;                  /* Gibbs.cpp:166 / 186-193                                             */
    const double* eel;            /* M + 1 expected effective lengths (WriteResults.h:24-53)             */
    const double* mw;             /* M + 1                                                               */
    int32_t n_genes;              /* m                                                                   */
    const int32_t* gene_start;    /* m + 1 (GroupInfo::spAt)                                             */
} rsem_b200_gibbs_params;

typedef struct rsem_b200_gibbs_out {
    /* per chain, concatenated in chain order: chain_samples[t] count vectors of M + 1 int32 each,
     * i.e. the lines of imd.countvectors<t> (Gibbs.cpp:257-262)                                 */
    int32_t* count_vectors;
    /* sums over ALL kept samples of all chains (Gibbs.cpp:325-345, 372-397)                     */
    double* sum_c;        /* M + 1 */
    double* sum_c2;       /* M + 1 */
    double* sum_tpm;      /* M + 1 */
    double* sum_fpkm;     /* M + 1 */
    double* sum_gene_c2;  /* n_genes */
} rsem_b200_gibbs_out;

int rsem_b200_gibbs_upload(rsem_b200_ctx* ctx, uint64_t N1, uint64_t E, int32_t M, const uint64_t* row_ptr,
                           const int32_t* sid, const double* conprb);
int rsem_b200_gibbs_run(rsem_b200_ctx* ctx, const rsem_b200_gibbs_params* params, rsem_b200_gibbs_out* out);

/* ---- instrumentation (bench / tests) ----------------------------------------------------------- */
/* number of kernels this library launched on ctx since creation */
int rsem_b200_launch_count(rsem_b200_ctx* ctx, uint64_t* launches);
/* device time (ms, CUDA events on the context's stream) spent in the E/M kernel (K2) launches and
 * their number since the last reset; used for the roofline figure.                              */
int rsem_b200_estep_timing(rsem_b200_ctx* ctx, double* total_ms, uint64_t* launches, int32_t reset);
/* enable per-launch event timing of K2 (off by default: it adds two event records per round)    */
int rsem_b200_set_profiling(rsem_b200_ctx* ctx, int32_t enabled);
/* select the E/M kernel variant: 0 = auto (equivalence-class layout for frozen-conprb rounds, row groups when
 * posteriors are written), 1 = CTA-staged tiles (TMA), 2 = direct (no smem staging), 3 = warp-pipelined tiles
 * (TMA, no CTA barriers), 4 = row groups on CTA-staged tiles (TMA, no CTA barriers, dynamic row batches),
 * 5 = equivalence-class layout, fail if the matrix is outside its limits                           */
int rsem_b200_set_estep_variant(rsem_b200_ctx* ctx, int32_t variant);

/* Equivalence-class layout the frozen-conprb rounds run on (derived from the hit matrix on first use; reads with
 * identical transcript lists - the reference has no counterpart, it walks HitContainer row by row, EM.cpp:199-244).
 * out[8] = { built (0/1), rows in the layout, rows handled by the long-row launch, segments, batches, tiles,
 *            staged doubles (conprb + ncpv), staged ids }.  Streamed bytes per round = 8 out[6] + 4 out[7] + 16 out[4]. */
int rsem_b200_class_layout_info(rsem_b200_ctx* ctx, uint64_t* out /* 8 */);
/* Busy time (ns, %globaltimer) of every persistent CTA in the latest class-layout E-step launch: the evidence for
 * the segmented reduction's load balance on skewed matrices (BASELINE configs[4]).  *n = CTAs written (<= cap).   */
int rsem_b200_estep_cta_times(rsem_b200_ctx* ctx, uint64_t* out_ns, int32_t cap, int32_t* n);

#ifdef __cplusplus
}
#endif
#endif /* RSEM_B200_H_ */
