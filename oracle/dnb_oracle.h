/* TEST INFRASTRUCTURE -- CPU oracle ("port"): a plain-C restatement of the DNAscent detect signal hot path.
 * Never shipped, never linked into the product; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors for this path, so the
 * restatement is pinned against outputs of the reference itself: oracle/_ref (the unmodified reference sources
 * compiled in the build container) on seeded synthetic reads, and against tests/golden/ fixtures generated from
 * oracle/_ref by tests/golden/make_golden.py.
 */
#ifndef DNB_ORACLE_H
#define DNB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNBO_K 9
#define DNBO_BANDWIDTH 100

typedef struct {
    uint32_t w1, w2;
    float thr1, thr2, peak_height;
} dnbo_detector_param; /* event_detection.h:10-25 defaults {3,6,1.4f,9.0f,0.2f} */

/* status codes of dnbo_normalise */
enum { DNBO_OK = 0, DNBO_QC_FAIL = 1, DNBO_SCALE_FAIL = 2, DNBO_UNDEFINED = 3 };

typedef struct {
    /* segmentation (event_detection.c:268-319) */
    size_t et_n;          /* scrappie event count */
    uint64_t *et_start;   /* [et_n] */
    float *et_length;     /* [et_n] */
    float *et_mean;       /* [et_n] */
    float *et_stdv;       /* [et_n] */
    /* r.events (event_handling.cpp:549-575) */
    size_t n_events;
    double *ev_mean;      /* [n_events] */
    uint32_t *ev_start;   /* [n_events+1] raw slice of event j = [ev_start[j], ev_start[j+1]) */
    /* ranks */
    size_t n_kmers, n_kmers_ref;
    uint32_t *rank_query, *rank_ref;
    /* scalings */
    double rough_shift, rough_scale, shift, scale, events_per_base;
    /* banded alignment */
    size_t n_bands;
    uint8_t *band_move;   /* [n_bands] 1 = right, 0 = down (bands 0,1: 0) */
    uint8_t *trace;       /* [n_bands*100] FROM_D=0 FROM_U=1 FROM_L=2 */
    float *last_col;      /* [n_events] bands score at (event, last k-mer) or -inf */
    double lp_skip, lp_stay, lp_step, lp_trim;
    int64_t fills;
    size_t n_align;       /* pairs BEFORE the QC clear */
    uint32_t *align_event, *align_kmer;
    size_t n_cleaned;
    double *cleaned_signal;
    uint32_t *cleaned_rank;
    double avg_log_emission;
    int spanned;
    int max_gap;
    float end_score;
    int status;
} dnbo_result;

void dnbo_result_free(dnbo_result *r);

/* a1-a5 */
size_t dnbo_detect_events(const float *raw, size_t n, dnbo_detector_param p, uint64_t *start, float *length, float *mean,
                          float *stdv, size_t cap);
/* a2 exposed for unit tests */
void dnbo_tstat(const float *raw, size_t n, uint32_t w, float *out);
/* a7 */
uint32_t dnbo_kmer2index(const char *kmer, unsigned k);
void dnbo_kmer_ranks(const char *seq, size_t len, uint32_t *out);
/* a8 */
int dnbo_quantile_scaling(const double *ev_mean, size_t n_events, const uint32_t *rank_ref, size_t n_ref,
                          const double *model_mean, double *shift, double *scale);
/* a9 */
float dnbo_log_probability_match(double ev_mean, double shift, double scale, double mu, double sigma);
/* a12 */
void dnbo_theil_sen(const double *sig, const uint32_t *ranks, size_t n, const double *model_mean, double shift,
                    double scale, double *out_shift, double *out_scale);
/* a13: the whole path.  model_stdv == NULL means the static 0.14 of the ONT table (useFitPoreModel=false). */
int dnbo_normalise(const float *raw, size_t n_raw, const char *query, size_t qlen, const char *ref, size_t rlen,
                   const int32_t *query_to_ref, const double *model_mean, const double *model_stdv, int keep_bands,
                   dnbo_result *out);
/* read loop of detect.cpp:852-876 over the port (CPU baseline when oracle/_ref cannot be built) */
double dnbo_bench_normalise(const float *const *raw, const size_t *n_raw, const char *const *query, const size_t *qlen,
                            const char *const *ref, const size_t *rlen, const int32_t *const *q2r, size_t n_reads,
                            const double *model_mean, int threads, int *failed);

/* a14 */
double dnbo_eexp(double x);
int dnbo_eln(double x, double *out); /* returns 1 for x<0 (reference throws NegativeLog) */
double dnbo_lnSum(double a, double b);
double dnbo_lnProd(double a, double b);
int dnbo_lnGreaterThan(double a, double b);
double dnbo_uniformPDF(double lb, double ub, double x);
double dnbo_normalPDF(double mu, double sigma, double x);
double dnbo_cauchyPDF(double loc, double scale, double x);
/* a15 */
double dnbo_sequence_probability(const double *obs, size_t n_obs, const char *seq, size_t seq_len, size_t window,
                                 int use_analogue, double shift, double scale, double events_per_base,
                                 size_t a_start, size_t a_end, const double *unl_mean, const double *unl_stdv,
                                 const double *ana_mean, const double *ana_stdv);
/* llAcrossRead (detect.cpp:393-574): returns #calls; pos = index on referenceSeqMappedTo (sequencing orientation) */
size_t dnbo_ll_across_read(const char *ref, size_t rlen, const int32_t *ref_to_query, int is_reverse,
                           const uint32_t *align_event, const uint32_t *align_kmer, size_t n_align,
                           const double *ev_mean, double shift, double scale, double events_per_base, unsigned window,
                           const double *unl_mean, const double *unl_stdv, const double *ana_mean,
                           const double *ana_stdv, uint32_t *pos, double *llr, size_t cap);

/* f1: builtinViterbi (alignment.cpp:193-516); path in order, type 0 = D, 1 = M, 2 = I; model_stdv NULL = 0.14 */
size_t dnbo_builtin_viterbi(const double *obs, size_t T, const char *seq, size_t seq_len, double shift, double scale,
                            double events_per_base, const double *model_mean, const double *model_stdv, double *score,
                            int32_t *idx, uint8_t *type, size_t cap);
/* f1: eventalign (alignment.cpp:547-744) as records (event, ref_pos, label 1 = M / 2 = I, indelScore) */
size_t dnbo_eventalign(const char *ref, size_t rlen, const int32_t *ref_to_query, const uint32_t *align_event,
                       const uint32_t *align_kmer, size_t n_align, const double *ev_mean, double shift, double scale,
                       double events_per_base, unsigned total_window, const double *model_mean, uint32_t *rec_event,
                       uint32_t *rec_refpos, uint8_t *rec_label, int32_t *rec_indel, size_t cap);

/* f2: the DNN input tensors (reads.h:305-372) from the eventalign records and the raw signal: signal [P][20],
 * core / residual k-mer indices, reference coordinates / indices, query indices, alignment quality; returns P.
 * called = sorted keys of r.refCoordToCalls (positions eventalign does not addSignal for, alignment.cpp:711). */
#define DNBO_RAWDEPTH 20
size_t dnbo_dnn_features(const char *ref, size_t rlen, const int32_t *ref_to_query, int is_reverse, uint32_t ref_start,
                         uint32_t ref_end, const uint32_t *rec_event, const uint32_t *rec_refpos,
                         const uint8_t *rec_label, const int32_t *rec_indel, size_t n_rec, const double *raw,
                         const uint32_t *event_start, double shift, double scale, const uint32_t *called,
                         size_t n_called, float *signal, float *core, float *residual, uint32_t *coords,
                         uint32_t *ref_index, uint32_t *query_index, int32_t *quality, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
