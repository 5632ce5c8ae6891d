/* dnascent_b200.h -- C ABI of libdnascent_b200.so
 *
 * B200-native (sm_100a CUDA) implementation of the per-read signal hot path of `DNAscent detect`
 * (MBoemo/DNAscent v4.1.1).  The reference has no plugin/FFI layer: the seam is the direct C++ call
 *     void normaliseEvents(DNAscent::read&, bool)            src/event_handling.h:13
 * made from the OpenMP read loops at src/detect.cpp:876, src/alignment.cpp:856, src/trainCNN.cpp:319,
 * plus   event_table detect_events(double*, size_t, detector_param)   src/scrappie/event_detection.h:35
 *        eexp/eln/lnSum/lnProd/lnGreaterThan/uniformPDF/normalPDF/cauchyPDF   src/probability.h:26-33
 *        sequenceProbability / llAcrossRead                  src/detect.h:119,121
 * This header is the batched, plain-C boundary those C++ symbols are re-implemented on (see
 * dnascent_b200/csrc/shim/ and INTEGRATION.md).  Plain pointers and sizes only, int error codes, no
 * exceptions cross the ABI.  There is NO CPU fallback: every entry point that computes needs a CUDA
 * device and fails with DNB_ERR_CUDA otherwise.
 */
#ifndef DNASCENT_B200_H
#define DNASCENT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNB_API __attribute__((visibility("default")))

/* ---- error codes (returned by every int function; 0 = success) ------------------------------ */
enum {
    DNB_OK = 0,
    DNB_ERR_ARG = 1,      /* NULL / out-of-range argument */
    DNB_ERR_CUDA = 2,     /* CUDA runtime error or no device (dnb_last_error() has the text) */
    DNB_ERR_NOMEM = 3,    /* host or device allocation failed */
    DNB_ERR_MODEL = 4,    /* required pore-model table not loaded */
    DNB_ERR_STATE = 5,    /* call order violated (e.g. result before wait) */
    DNB_ERR_NEGATIVE_LOG = 6 /* dnb_eln(x<0): the reference throws NegativeLog (src/probability.cpp:45) */
};

/* ---- per-read status (dnb_read_result.status) ------------------------------------------------
 * The reference signals failure by leaving r.eventAlignment empty (src/detect.cpp:879); the status says why. */
enum {
    DNB_READ_OK = 0,
    DNB_READ_QC_FAIL = 1,      /* avg emission / spanned / max_gap / <1000 cleaned points  (event_handling.cpp:433-441) */
    DNB_READ_SCALE_FAIL = 2,   /* Theil-Sen slope == 0 -> scalings {-1,-1}               (event_handling.cpp:90-95,604) */
    DNB_READ_UNDEFINED = 3,    /* inputs for which the reference itself is undefined (no peak found, no event,
                                  query shorter than k+1, end cell outside the band: it would index out of range) */
    DNB_READ_OVERFLOW = 4      /* more events than the per-read device capacity (resubmit with a larger
                                  dnb_config.event_capacity_per_sample) */
};

/* which table dnb_load_model fills; src/config.h:39 (pore_model, unlabelled_model, analogue_model) */
enum { DNB_MODEL_PORE = 0, DNB_MODEL_UNLABELLED = 1, DNB_MODEL_ANALOGUE = 2 };

#define DNB_KMER_LEN 9
#define DNB_N_KMERS 262144 /* 4^9 */
#define DNB_MAX_DEVICES 16

/* dnb_config.result_format: what dnb_submit / dnb_batch_fetch bring back over PCIe (see dnb_read_result) */
enum { DNB_RESULT_DENSE = 0, DNB_RESULT_COMPACT = 1 };

typedef struct dnb_ctx dnb_ctx;
typedef struct dnb_batch dnb_batch;

/* Everything the reference keeps in compile-time constants / Global_Config (src/config.h:41-63,
 * src/scrappie/event_detection.h:19-25). dnb_default_config() fills the R10.4.1 DNA values. */
typedef struct {
    int device;                       /* CUDA ordinal */
    uint32_t window_length1;          /* 3     detector_param */
    uint32_t window_length2;          /* 6 */
    float threshold1;                 /* 1.4f */
    float threshold2;                 /* 9.0f */
    float peak_height;                /* 0.2f */
    double min_average_log_emission;  /* -2.0  AdaptiveBanded_Params */
    int max_gap_threshold;            /* 5 */
    int bandwidth;                    /* 100 (only value supported) */
    int use_fit_pore_model;           /* the `useFitPoreModel` argument of normaliseEvents; all reference callers pass false */
    float event_capacity_per_sample;  /* device event slots per raw sample (default 0.40; observed ~0.2) */
    int keep_debug;                   /* also return rough scalings' inputs: cleaned (signal,rank) vectors */
    size_t workspace_bytes;           /* cap for the idle device blocks the context keeps cached between batches
                                         (0 = no cap; dnb_trim() releases them on request) */
    int result_format;                /* DNB_RESULT_DENSE (default) or DNB_RESULT_COMPACT */
    int n_devices;                    /* 0 or 1: `device` only.  > 1: one context drives devices[0 .. n_devices): tables are
                                         replicated, every dnb_submit / dnb_submit_chain / dnb_batch_upload goes to the
                                         device with the least work in flight (reads are independent: no collective) */
    int devices[DNB_MAX_DEVICES];
} dnb_config;

/* queryToRef[q_start + i] = r_start + (stride ? i : 0) for i < len (stride 1: aligned bases; stride 0: the entries
 * parseCigar gives insertions / soft clips, src/htsInterface.cpp:143-152).  Query positions no run covers have no
 * entry.  Later runs overwrite earlier ones where they overlap. */
typedef struct dnb_q2r_run {
    uint32_t q_start, len;
    int32_t r_start;
    int32_t stride;               /* 0 or 1 */
} dnb_q2r_run;

/* One read, exactly the fields normaliseEvents reads from DNAscent::read (src/reads.h:178-208):
 * raw (float32-exact pA, src/pod5.cpp:60), basecall, referenceSeqMappedTo, queryToRef. */
typedef struct {
    const float *raw_pA;          /* n_samples values, or NULL when raw_dac is given */
    const int16_t *raw_dac;       /* optional: int16 DAC; pA = ((float)dac + dac_offset) * dac_scale (pod5.cpp:60) */
    float dac_offset, dac_scale;
    uint64_t n_samples;
    const char *query;            /* r.basecall, sequencing orientation */
    uint32_t query_len;
    const char *ref;              /* r.referenceSeqMappedTo, sequencing orientation */
    uint32_t ref_len;
    const int32_t *query_to_ref;  /* dense r.queryToRef: query_len entries, -1 = no entry; or NULL when q2r_runs is given */
    const struct dnb_q2r_run *q2r_runs;   /* r.queryToRef as runs (what parseCigar produces, src/htsInterface.cpp:59-157): */
    uint32_t n_q2r_runs;                  /* 16 B per CIGAR operation instead of 4 B per base over PCIe */
} dnb_read_desc;

typedef struct {
    int status;                   /* DNB_READ_* */
    uint32_t et_n;                /* scrappie event count (event_table.n) */
    uint32_t n_events;            /* r.events.size() */
    const uint32_t *event_start;  /* [n_events+1]: r.events[j].raw == raw[event_start[j] .. event_start[j+1]) */
    const float *event_mean;      /* [n_events]:   r.events[j].mean (float32-exact; [0] is 0.0, quirk Q1) */
    uint32_t n_align;             /* r.eventAlignment.size(); 0 unless status == DNB_READ_OK */
    const uint32_t *align_pairs;  /* [2*n_align] interleaved (event_idx, kmer_idx) == vector<pair<unsigned,unsigned>> layout */
    double shift, scale, events_per_base;  /* r.scalings (Theil-Sen refined) */
    double rough_shift, rough_scale;       /* quantile scaling used by the alignment (event_handling.cpp:595) */
    double avg_log_emission;               /* r.alignmentQCs */
    int spanned;
    int max_gap;
    uint32_t n_cleaned;           /* keep_debug only */
    const double *cleaned_signal;
    const uint32_t *cleaned_rank;
    /* ---- DNB_RESULT_COMPACT: event_start and align_pairs above are NULL; the same information comes back as
     * 1 B per event and 2 bits per alignment step (dense: 4 B and 8 B), dnb_expand_* rebuild the dense arrays -------- */
    uint32_t event_first;            /* event_start[0] */
    const uint8_t *event_len8;       /* [n_events]: event_start[j+1] - event_start[j]; 255 = the next entry of ... */
    const uint32_t *event_len_escape;/* ... this list (lengths >= 255 samples, in event order) */
    uint32_t n_event_len_escape;
    uint32_t align_first[2];         /* eventAlignment[0] = (event, kmer) */
    const uint8_t *align_steps;      /* n_align - 1 steps, 2 bits each, step t at bits 2*(t&3) of byte t>>2:
                                        0: (event+1, kmer+1)   1: (event+1, kmer)   2: (event, kmer+1)  -- the three
                                        moves of the backtrace (src/event_handling.cpp:160-162) read forwards */
} dnb_read_result;

/* scrappie event table entry (src/scrappie/scrappie_structures.h:8-15) for the detect_events drop-in */
typedef struct {
    uint64_t start;
    float length;
    float mean;
    float stdv;
    int pos;
    int state;
} dnb_event_t;

/* ---- lifecycle ---------------------------------------------------------------------------- */
DNB_API void dnb_default_config(dnb_config *cfg);
DNB_API int dnb_create(dnb_ctx **ctx, const dnb_config *cfg);
DNB_API void dnb_destroy(dnb_ctx *ctx);
DNB_API const char *dnb_strerror(int code);
DNB_API const char *dnb_last_error(void); /* thread-local detail of the last DNB_ERR_CUDA */
/* replaces import_poreModel_staticStdv / import_poreModel_fitStdv results (src/data_IO.cpp:144-242):
 * mean/stdv indexed by kmer2index (A=0,T=1,G=2,C=3, first base most significant); n must be 4^9. */
DNB_API int dnb_load_model(dnb_ctx *ctx, int which, const double *mean, const double *stdv, size_t n);

/* ---- the hot path: batched normaliseEvents (src/event_handling.cpp:544-607) ---------------- */
/* Copies the reads to the device (straight from the caller's buffers when they are page-locked, through pinned staging
 * otherwise), runs segmentation -> quantile scaling -> banded alignment -> backtrace/QC -> Theil-Sen and copies the
 * results back; returns when they are on the host (dnb_wait is then a no-op kept for the split form).
 * Thread-safe: concurrent callers (the OpenMP read loop of detect.cpp:852) form a software pipeline -- one batch
 * copying in, two computing, one copying out -- which is how copies and kernels overlap. */
DNB_API int dnb_submit(dnb_ctx *ctx, const dnb_read_desc *reads, size_t n_reads, dnb_batch **batch);
DNB_API int dnb_wait(dnb_batch *batch);
DNB_API int dnb_result(dnb_batch *batch, size_t i, dnb_read_result *out);
DNB_API void dnb_release(dnb_batch *batch);
/* Rebuild the dense arrays of a DNB_RESULT_COMPACT result in caller memory (e.g. straight into the std::vector the
 * reference keeps them in): event_start[n_events + 1], align_pairs[2 * n_align].  Dense results are copied. */
DNB_API int dnb_expand_events(const dnb_read_result *res, uint32_t *event_start);
DNB_API int dnb_expand_alignment(const dnb_read_result *res, uint32_t *align_pairs);

/* ---- caller-owned pinned memory: zero-staging ingest ----------------------------------------------------------
 * dnb_submit copies a read's signal to the device straight from the caller's buffer when that buffer is page-locked
 * (registered here, allocated here, or by the caller's own cudaHostAlloc / cudaHostRegister); pageable buffers are
 * first packed into the library's pinned staging area (one extra pass over host memory).  A POD5 loader that
 * decompresses into dnb_host_alloc'ed memory therefore feeds the GPUs without touching the samples again. */
DNB_API int dnb_host_register(void *p, size_t bytes);
DNB_API int dnb_host_unregister(void *p);
DNB_API int dnb_host_alloc(void **p, size_t bytes);
DNB_API void dnb_host_free(void *p);
/* give the device and pinned blocks the context keeps cached between batches back to the driver */
DNB_API int dnb_trim(dnb_ctx *ctx);

/* Split form of dnb_submit for callers that keep inputs resident in HBM (bench `value` leg):
 * upload once, run the device pipeline any number of times, fetch results when wanted. */
DNB_API int dnb_batch_upload(dnb_ctx *ctx, const dnb_read_desc *reads, size_t n_reads, dnb_batch **batch);
DNB_API int dnb_batch_run(dnb_batch *batch);     /* device pipeline only; blocks until done */
DNB_API int dnb_batch_fetch(dnb_batch *batch);   /* device -> pinned host results */
/* returns the batch's device workspace (~27 B/sample) to the pool; the inputs stay resident, results already
 * fetched stay valid.  dnb_submit does this itself once the results are on the host. */
DNB_API int dnb_batch_drop_workspace(dnb_batch *batch);
/* device-time breakdown of the last dnb_batch_run, milliseconds, measured with CUDA events on the
 * pipeline stream: [0]=segmentation [1]=ranks+scaling+prep [2]=banded DP [3]=backtrace+QC [4]=Theil-Sen [5]=total
 * (first to last kernel, host step included); host wall clock: [6]=the mid-pipeline host step (transition constants,
 * workspace sizing) [7]=the whole dnb_batch_run call;
 * counts: [0]=samples [1]=events [2]=k-mers [3]=bands [4]=DP cells [5]=kernel launches
 *         [6]=reads the tiled segmentation handed to its serial kernel [7]=reads with status != DNB_READ_OK */
DNB_API int dnb_batch_timings(dnb_batch *batch, double ms[8], uint64_t counts[8]);
/* the segmentation part of ms[0], by kernel (CUDA events): [0] exact (sum, sumsq) checkpoints (scan or serial chain)
 * [1] tiles (t-statistics + peak detectors)  [2] stitch + events + serial redo of flagged reads */
DNB_API int dnb_batch_seg_timings(dnb_batch *batch, double ms[3]);
/* Host-side accounting of the batch pipeline (process-wide, all contexts): wall seconds spent in every phase of
 * upload / run / fetch, summed over the calling threads since the last reset -- where the host time of dnb_submit goes
 * (waiting for a pipeline slot, packing, enqueueing, waiting for the GPU).  counts: [0] batches uploaded [1] of them
 * with the signal DMA'd straight from page-locked caller memory [2] cudaMalloc calls [3] cudaMallocHost calls (both
 * should be 0 in steady state: the context caches its blocks). */
#define DNB_N_HOST_PHASES 17
DNB_API int dnb_host_stats(int reset, double seconds[DNB_N_HOST_PHASES], uint64_t counts[4]);
DNB_API const char *dnb_host_phase_name(int phase);
/* device ordinal the batch was dealt to (contexts created with n_devices > 1) */
DNB_API int dnb_batch_device(dnb_batch *batch);
/* PCIe payload of this batch, counted from the copies made: host->device bytes of the upload (signal, sequences,
 * queryToRef, offset tables) and device->host bytes of the fetch (dense events, alignment pairs, per-read scalars) */
DNB_API int dnb_batch_io_bytes(dnb_batch *batch, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* ---- detect_events drop-in (src/scrappie/event_detection.h:35) ------------------------------ */
/* raw_pA: n float32-exact samples.  events: caller array of capacity cap; *n_events receives event_table.n. */
DNB_API int dnb_detect_events(dnb_ctx *ctx, const float *raw_pA, size_t n, dnb_event_t *events, size_t cap,
                              size_t *n_events);

/* ---- probability.cpp drop-ins (src/probability.h:26-33); NaN == log(0) convention ------------- */
DNB_API double dnb_eexp(double x);
DNB_API int dnb_eln(double x, double *out); /* DNB_ERR_NEGATIVE_LOG where the reference throws */
DNB_API double dnb_lnSum(double ln_x, double ln_y);
DNB_API double dnb_lnProd(double ln_x, double ln_y);
DNB_API int dnb_lnGreaterThan(double ln_x, double ln_y);
DNB_API double dnb_uniformPDF(double lb, double ub, double x);
DNB_API double dnb_normalPDF(double mu, double sigma, double x);
DNB_API double dnb_cauchyPDF(double loc, double scale, double x);

/* ---- Theil-Sen refinement alone: estimateScaling_theilSen (src/event_handling.cpp:24-110) ------ */
/* Read i: cleaned (signal, k-mer rank) vectors signals/ranks[off[i] .. off[i+1]) as normaliseEvents builds them
 * (event_handling.cpp:386-393) and the rough scalings of the quantile fit; shift/scale receive what the reference
 * function returns with useFitPoreModel == false (DNB_MODEL_PORE levels): the rough values for fewer than 1000
 * points, (-1, -1) for a zero median slope.  The same kernels dnb_submit runs, including the path for 0/0 slopes
 * (a NaN among the slopes the reference std::sorts: the result is the one libstdc++'s introsort gives). */
DNB_API int dnb_theil_sen_batch(dnb_ctx *ctx, const double *signals, const uint32_t *ranks, const uint64_t *off,
                                size_t n_reads, const double *rough_shift, const double *rough_scale, double *shift,
                                double *scale);

/* ---- analogue likelihood: batched sequenceProbability (src/detect.cpp:235-378) ---------------- */
/* One "site" = one call of sequenceProbability: observations obs[obs_off[s] .. obs_off[s+1]) (event means, pA),
 * a (2*window + 9)-base snippet at seq + s*(2*window+9), per-site scalings.  Computes both the analogue pass
 * (useBrdU=true, BrdUStart/End = window -/+ 4) and the thymidine pass, as llAcrossRead does (detect.cpp:546-548).
 * BrdUStart / BrdUEnd are NOT parameters here: window -/+ 4 is what the reference's only caller passes (detect.cpp:544-545),
 * and the C++ shim's sequenceProbability throws std::invalid_argument for any other span instead of scoring it differently.
 * out_analogue / out_thymidine: log forward probabilities (NaN == log 0); LLR = analogue - thymidine. */
DNB_API int dnb_sequence_probability_batch(dnb_ctx *ctx, const double *obs, const uint64_t *obs_off, const char *seq,
                                           const double *shift, const double *scale, const double *events_per_base,
                                           size_t n_sites, uint32_t window, double *out_analogue,
                                           double *out_thymidine);

/* ---- eventalign: windowed Viterbi re-alignment (src/alignment.h:22, src/alignment.cpp:193-516, 547-744) ---- */
/* The stage that follows normaliseEvents in the read loop (src/detect.cpp:888, src/alignment.cpp:866,
 * src/trainCNN.cpp:328): per read a serial chain of ~50-base reference windows, each re-aligned by builtinViterbi.
 * One descriptor = the fields eventalign reads from DNAscent::read (src/reads.h:178-208). */
typedef struct {
    const char *ref;               /* r.referenceSeqMappedTo */
    uint32_t ref_len;
    const int32_t *ref_to_query;   /* dense r.refToQuery, ref_len entries; an absent key reads as 0 (std::map::operator[]) */
    const uint32_t *align_pairs;   /* r.eventAlignment, interleaved (event_idx, kmer_idx) */
    uint32_t n_align;
    const float *event_mean;       /* r.events[j].mean (float32-exact), n_events entries */
    uint32_t n_events;
    double shift, scale, events_per_base;   /* r.scalings */
} dnb_eventalign_desc;

/* One record per event the reference prints lines for (alignment.cpp:676-736): the event (index into r.events), the
 * position of its k-mer on referenceSeqMappedTo (reference_index + pos), the window's indelScore (alignment.cpp:638)
 * and the state label.  Printed coordinate: refStart + ref_pos + 4 (fwd) / refEnd - ref_pos - 5 (rev). */
enum { DNB_EA_MATCH = 1, DNB_EA_INSERTION = 2 };
typedef struct dnb_eventalign_rec {
    uint32_t event;
    uint32_t ref_pos;
    int32_t indel_score;
    uint32_t label;                /* DNB_EA_MATCH / DNB_EA_INSERTION */
} dnb_eventalign_rec;

/* recs: caller array; read i owns recs[rec_off[i] .. rec_off[i+1]) (capacity; n_align + 64 always suffices on
 * well-formed alignments); n_recs[i] receives the count, status[i] a DNB_READ_* code (DNB_READ_UNDEFINED where the
 * reference itself throws or indexes out of range: events_per_base <= 1, ref_len < 9; DNB_READ_OVERFLOW when
 * the record capacity or the per-window event capacity (4096) is exceeded).  Needs DNB_MODEL_PORE. */
DNB_API int dnb_eventalign_batch(dnb_ctx *ctx, const dnb_eventalign_desc *reads, size_t n_reads, uint32_t window,
                                 dnb_eventalign_rec *recs, const uint64_t *rec_off, uint32_t *n_recs, int *status);
/* device time (ms, CUDA events around the kernel) of the last dnb_eventalign_batch on this thread */
DNB_API double dnb_eventalign_last_kernel_ms(void);

/* ---- DNN input tensors built on the device (src/reads.h:147-172, 288-452; consumer src/detect.cpp:586-649) ---- */
/* eventalign's side effect in the reference is r.addSignal() per raw sample of every match-state event
 * (alignment.cpp:706-725); runCNN then turns r.refCoordToAP into the tensors it feeds TensorFlow.  This entry runs
 * eventalign and builds those tensors in one device pass (records never leave the GPU unless asked for), in the
 * identical layout: row o of every output == the o-th element the reference's makeSignalTensor /
 * makeCoreSequenceTensor / makeResidualSequenceTensor / getReferenceCoords / getReferenceIndices / getQueryIndices /
 * getAlignmentQuality produce (strand-dependent iteration order included). */
#define DNB_RAWDEPTH 20 /* src/reads.h:12 */

/* what addSignal / the tensor builders read from DNAscent::read besides dnb_eventalign_desc */
typedef struct {
    const float *raw_pA;           /* r.raw (float32-exact pA), or NULL when raw_dac is given */
    const int16_t *raw_dac;        /* int16 DAC; pA = ((float)dac + dac_offset) * dac_scale (src/pod5.cpp:60) */
    float dac_offset, dac_scale;
    uint64_t n_samples;
    const uint32_t *event_start;   /* [n_events + 1] as dnb_read_result.event_start: r.events[j].raw = raw[start[j], start[j+1]) */
    int is_reverse;                /* r.isReverse */
    uint32_t ref_start, ref_end;   /* r.refStart, r.refEnd */
    const uint32_t *called;        /* ascending keys of r.refCoordToCalls (alignment.cpp:711), may be NULL */
    uint32_t n_called;
} dnb_feature_desc;

/* flat outputs; read i owns rows [pos_off[i], pos_off[i] + n_pos[i]) of each */
typedef struct {
    float *signal;                 /* [rows][DNB_RAWDEPTH]  makeSignalTensor: scaled samples, zero padded */
    float *core;                   /* [rows] getCoreIndex() = 1 + base-4 rank of k-mer bases 2..6 */
    float *residual;               /* [rows] getResidualIndex() = 1 + rank of bases 0,1,7,8 */
    uint32_t *coords;              /* [rows] getReferenceCoords (genome coordinate) */
    uint32_t *ref_index;           /* [rows] getReferenceIndices (index on referenceSeqMappedTo) */
    uint32_t *query_index;         /* [rows] getQueryIndices */
    int32_t *quality;              /* [rows] getAlignmentQuality (the window's indelScore) */
} dnb_feature_tensors;

/* reads / recs / rec_off / n_recs / status as dnb_eventalign_batch (recs may be NULL: records stay on the device).
 * pos_off: [n_reads+1] row capacity per read (ref_len - 8 always suffices); n_pos[i] receives the row count.
 * status additionally reports DNB_READ_OVERFLOW when the row capacity is too small. */
DNB_API int dnb_eventalign_features_batch(dnb_ctx *ctx, const dnb_eventalign_desc *reads, const dnb_feature_desc *feats,
                                          size_t n_reads, uint32_t window, dnb_eventalign_rec *recs,
                                          const uint64_t *rec_off, uint32_t *n_recs, int *status,
                                          const dnb_feature_tensors *out, const uint64_t *pos_off, uint32_t *n_pos);
/* device time (ms) of the feature kernel of the last dnb_eventalign_features_batch on this thread */
DNB_API double dnb_features_last_kernel_ms(void);

/* ---- resident form: eventalign + DNN input tensors on a batch that dnb_batch_run has processed ----------------- */
/* The read loop of detect.cpp:876-888 as one device-resident chain: the signal, the events, the alignment and the
 * scalings of normaliseEvents are already in HBM, so eventalign and the tensor builder read them in place -- the only
 * additional host->device bytes are the fields below, the only device->host bytes the tensors.  Split form only
 * (dnb_batch_upload / dnb_batch_run, workspace not dropped); runs dnb_batch_fetch itself if that has not happened. */
typedef struct {
    const int32_t *ref_to_query;   /* dense r.refToQuery, ref_len entries; an absent key reads as 0 */
    int is_reverse;                /* r.isReverse */
    uint32_t ref_start, ref_end;   /* r.refStart, r.refEnd */
    const uint32_t *called;        /* ascending keys of r.refCoordToCalls, may be NULL */
    uint32_t n_called;
} dnb_read_extra;

typedef struct {
    int status;                    /* DNB_READ_*: normaliseEvents' status, or what eventalign / the tensor builder found */
    uint32_t n_pos;                /* rows */
    const float *signal;           /* [n_pos][DNB_RAWDEPTH] */
    const float *core, *residual;  /* [n_pos] */
    const uint32_t *coords, *ref_index, *query_index;
    const int32_t *quality;
    uint32_t n_recs;               /* eventalign records (0 and NULL unless want_records) */
    const dnb_eventalign_rec *recs;
} dnb_feature_result;

DNB_API int dnb_batch_eventalign_features(dnb_batch *batch, const dnb_read_extra *extra, uint32_t window,
                                          int want_records);
/* dnb_submit followed by dnb_batch_eventalign_features in one call: host buffers in, normaliseEvents results
 * (dnb_result) and tensors (dnb_batch_feature_result) out, workspace returned to the pool.  Thread-safe and staged like
 * dnb_submit, so concurrent callers overlap packing, copies and the two compute stages. */
DNB_API int dnb_submit_chain(dnb_ctx *ctx, const dnb_read_desc *reads, const dnb_read_extra *extra, size_t n_reads,
                             uint32_t window, int want_records, dnb_batch **batch);
/* pointers stay valid until dnb_release / the next dnb_batch_run */
DNB_API int dnb_batch_feature_result(dnb_batch *batch, size_t i, dnb_feature_result *out);
/* ms: [0] eventalign kernel, [1] feature kernel (CUDA events on the batch's stream); bytes: [0] host->device, [1] device->host */
DNB_API int dnb_batch_stage2_timings(dnb_batch *batch, double ms[2], uint64_t bytes[2]);

/* ---- resident analogue stage: llAcrossRead on what dnb_batch_run left in HBM (src/detect.cpp:393-574) --------------
 * `detect --HMM` calls llAcrossRead(r, 12) right after normaliseEvents (detect.cpp:885).  Here the T positions of
 * referenceSeqMappedTo, the events aligned to each site's window (readHead scan included) and both forward passes per
 * site are done on the device from the resident events / alignment / scalings: the only additional host->device bytes
 * are dnb_read_extra.ref_to_query and is_reverse, the only device->host bytes 24 B per candidate site.
 * Needs DNB_MODEL_UNLABELLED and DNB_MODEL_ANALOGUE (BrdU or EdU table: whichever was loaded). */
typedef struct {
    int status;                  /* normaliseEvents' DNB_READ_* status; a failed read has no sites */
    uint32_t n_sites;            /* T positions visited, in llAcrossRead's order (descending posOnRef for reverse reads) */
    const uint32_t *pos_on_ref;  /* [n_sites] posOnRef */
    const uint32_t *n_events;    /* [n_sites] events of the site's snippet; 0 = no call was made for this site
                                    (undefined snippet, or fewer than 2*window - 9 events: the `continue`s at :442, :515) */
    const double *log_analogue;  /* [n_sites] sequenceProbability(useBrdU = true)  (NaN == log 0), valid where n_events > 0 */
    const double *log_thymidine; /* [n_sites] sequenceProbability(useBrdU = false); LLR = analogue - thymidine (:546-548) */
} dnb_analogue_result;
DNB_API int dnb_batch_analogue_llr(dnb_batch *batch, const dnb_read_extra *extra, uint32_t window);
DNB_API int dnb_batch_analogue_result(dnb_batch *batch, size_t i, dnb_analogue_result *out);
/* dnb_submit followed by dnb_batch_analogue_llr in one pipelined call (the --HMM read loop body, detect.cpp:876-885) */
DNB_API int dnb_submit_llr(dnb_ctx *ctx, const dnb_read_desc *reads, const dnb_read_extra *extra, size_t n_reads,
                           uint32_t window, dnb_batch **batch);
/* ms: [0] site gathering kernel, [1] forward-pass kernel (CUDA events); counts: [0] candidate sites [1] calls made
 * [2] observations consumed by the forward passes (per pass) [3] device->host bytes */
DNB_API int dnb_batch_analogue_timings(dnb_batch *batch, double ms[2], uint64_t counts[4]);

/* ---- int16 ingest: the Dorado signal slice of pod5_getSignal (src/pod5.cpp:56-93, tags parsed at src/reads.h:221-253) -- */
/* The reference converts the whole POD5 record to pA and then erases what Dorado trimmed or what belongs to the
 * sibling of a split read.  Here the int16 DAC samples go to the device as they are (dnb_read_desc.raw_dac + the
 * record's calibration; pA is formed in registers with pod5.cpp:60's float expression), so trimming is a pointer
 * offset: only the slice is staged and copied.  This computes the slice [*first, *first + *count) of the record's
 * n_total samples exactly as the two vector::erase calls do:
 *   signal_length  r.signalLength (BAM tag ns), <= 0 when absent: no slicing (pod5.cpp:76)
 *   signal_trim    r.signalTrim (ts);  signal_start_coord  r.signalStartCoord (sp);
 *   is_split       r.readID != r.readID_fetch (a parent id, tag pi, was present)
 * Returns DNB_ERR_ARG where the reference's erase calls are undefined (start > end, or end past the record), and for an
 * empty record (the reference exits, pod5.cpp:64-73). */
DNB_API int dnb_dorado_slice(uint64_t n_total, int64_t signal_length, int64_t signal_trim, int64_t signal_start_coord,
                             int is_split, uint64_t *first, uint64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* DNASCENT_B200_H */
