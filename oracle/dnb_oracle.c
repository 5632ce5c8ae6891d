/* TEST INFRASTRUCTURE -- CPU oracle ("port").  See dnb_oracle.h for the rules and the parity status (PINNED
 * against oracle/_ref = the unmodified reference, and against tests/golden/).
 *
 * Plain-C restatement of the `DNAscent detect` signal hot path of MBoemo/DNAscent v4.1.1.  Every function
 * names the reference lines it follows (paths relative to /root/reference).  The arithmetic is transcribed
 * type-for-type (which operands are float, which are double, where values are rounded) because event
 * boundaries and alignment paths have to come out bit-identical; build with -ffp-contract=off.
 */
#include "dnb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include <stdio.h>

#define W DNBO_BANDWIDTH
#define KLEN DNBO_K

/* ------------------------------------------------------------------------------------------------
 * a1-a5  scrappie event detection          src/scrappie/event_detection.c
 * ---------------------------------------------------------------------------------------------- */

/* event_detection.c:35-48 -- strictly sequential double prefix sums (rounding order is part of the result) */
static void prefix_sums(const float *raw, size_t n, double *sum, double *sumsq) {
    sum[0] = 0.0;
    sumsq[0] = 0.0;
    for (size_t i = 0; i < n; i++) {
        double x = (double)raw[i];
        sum[i + 1] = sum[i] + x;
        sumsq[i + 1] = sumsq[i] + x * x;
    }
}

/* event_detection.c:60-115 */
static void tstat_from_sums(const double *sum, const double *sumsq, size_t n, size_t w, float *t) {
    const float wf = (float)w;
    memset(t, 0, n * sizeof(float));
    if (n < 2 * w || w < 2) return; /* :78-83 */
    for (size_t i = w; i <= n - w; i++) {                 /* :89, inclusive upper bound */
        double sum1 = sum[i], sumsq1 = sumsq[i];
        if (i > w) {
            sum1 -= sum[i - w];
            sumsq1 -= sumsq[i - w];
        }
        float sum2 = (float)(sum[i + w] - sum[i]);
        float sumsq2 = (float)(sumsq[i + w] - sumsq[i]);
        float mean1 = (float)(sum1 / (double)wf);
        float mean2 = sum2 / wf;
        float m1sq = mean1 * mean1;
        float q2 = sumsq2 / wf;
        float m2sq = mean2 * mean2;
        float cv = (float)(((sumsq1 / (double)wf - (double)m1sq) + (double)q2) - (double)m2sq);
        cv = fmaxf(cv, FLT_MIN);
        float dm = mean2 - mean1;
        float r = cv / wf;
        t[i] = (float)(fabs((double)dm) / sqrt((double)r));
    }
}

void dnbo_tstat(const float *raw, size_t n, uint32_t w, float *out) {
    double *sum = (double *)malloc((n + 1) * sizeof(double)), *sumsq = (double *)malloc((n + 1) * sizeof(double));
    prefix_sums(raw, n, sum, sumsq);
    tstat_from_sums(sum, sumsq, n, w, out);
    free(sum);
    free(sumsq);
}

typedef struct {
    const float *signal;
    float threshold;
    size_t window;
    size_t masked_to;
    int peak_pos;       /* -1 = none */
    float peak_value;
    int valid_peak;
} detector_t;

/* event_detection.c:122-198 -- short detector first, then long, at every sample */
static size_t peak_fsm(detector_t *ds, detector_t *dl, size_t n, float peak_height, size_t *peaks) {
    detector_t *det[2] = {ds, dl};
    size_t npk = 0;
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 2; k++) {
            detector_t *d = det[k];
            if (d->masked_to >= i) continue;                       /* :140 */
            float cur = d->signal[i];
            if (d->peak_pos == -1) {                               /* :146 */
                if (cur < d->peak_value) {
                    d->peak_value = cur;
                } else if (cur - d->peak_value > peak_height) {
                    d->peak_value = cur;
                    d->peak_pos = (int)i;
                }
            } else {
                if (cur > d->peak_value) {                         /* :159 */
                    d->peak_value = cur;
                    d->peak_pos = (int)i;
                }
                if (d == ds && d->peak_value > d->threshold) {     /* :165-176 short dominates long */
                    dl->masked_to = (size_t)d->peak_pos + d->window;
                    dl->peak_pos = -1;
                    dl->peak_value = FLT_MAX;
                    dl->valid_peak = 0;
                }
                if (d->peak_value - cur > peak_height && d->peak_value > d->threshold) d->valid_peak = 1; /* :178 */
                if (d->valid_peak && (i - (size_t)d->peak_pos) > d->window / 2) {  /* :183 */
                    peaks[npk++] = (size_t)d->peak_pos;
                    d->peak_pos = -1;
                    d->peak_value = cur;
                    d->valid_peak = 0;
                }
            }
        }
    }
    return npk;
}

/* event_detection.c:213-232 */
static void make_event(size_t start, size_t end, const double *sum, const double *sumsq, uint64_t *o_start,
                       float *o_len, float *o_mean, float *o_stdv) {
    float length = (float)(end - start);
    float mean = (float)(sum[end] - sum[start]) / length;
    float dsq = (float)(sumsq[end] - sumsq[start]);
    float var = dsq / length - mean * mean;
    *o_start = (uint64_t)start;
    *o_len = length;
    *o_mean = mean;
    *o_stdv = sqrtf(fmaxf(var, 0.0f));
}

/* event_detection.c:268-319 (+ create_events :234-266).  Returns et.n, 0 when the reference's behaviour is
 * undefined (no peak: it indexes peaks[n-2] with n == 1). */
size_t dnbo_detect_events(const float *raw, size_t n, dnbo_detector_param p, uint64_t *start, float *length, float *mean,
                          float *stdv, size_t cap) {
    if (n == 0) return 0;
    double *sum = (double *)malloc((n + 1) * sizeof(double)), *sumsq = (double *)malloc((n + 1) * sizeof(double));
    float *t1 = (float *)malloc(n * sizeof(float)), *t2 = (float *)malloc(n * sizeof(float));
    size_t *peaks = (size_t *)calloc(n, sizeof(size_t));
    prefix_sums(raw, n, sum, sumsq);
    tstat_from_sums(sum, sumsq, n, p.w1, t1);
    tstat_from_sums(sum, sumsq, n, p.w2, t2);
    detector_t ds = {t1, p.thr1, p.w1, 0, -1, FLT_MAX, 0}, dl = {t2, p.thr2, p.w2, 0, -1, FLT_MAX, 0};
    peak_fsm(&ds, &dl, n, p.peak_height, peaks);
    size_t ne = 1; /* :243-247 counts over the whole zero-padded array */
    for (size_t i = 0; i < n; i++)
        if (peaks[i] > 0 && peaks[i] < n) ne++;
    if (ne < 2 || ne > cap) {
        ne = 0;
    } else {
        float s_dummy;
        make_event(0, peaks[0], sum, sumsq, &start[0], &length[0], &mean[0], stdv ? &stdv[0] : &s_dummy);
        for (size_t e = 1; e + 1 < ne; e++)
            make_event(peaks[e - 1], peaks[e], sum, sumsq, &start[e], &length[e], &mean[e], stdv ? &stdv[e] : &s_dummy);
        make_event(peaks[ne - 2], n, sum, sumsq, &start[ne - 1], &length[ne - 1], &mean[ne - 1],
                   stdv ? &stdv[ne - 1] : &s_dummy);
    }
    free(sum); free(sumsq); free(t1); free(t2); free(peaks);
    return ne;
}

/* ------------------------------------------------------------------------------------------------
 * a7  k-mer ranks                           src/data_IO.cpp:129-141, src/event_handling.cpp:578-592
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t base_code(char c) { /* A=0 T=1 G=2 C=3, anything else 0 (std::map::operator[] default) */
    return c == 'T' ? 1u : c == 'G' ? 2u : c == 'C' ? 3u : 0u;
}
uint32_t dnbo_kmer2index(const char *kmer, unsigned k) {
    uint32_t r = 0;
    for (unsigned i = 0; i < k; i++) r = r * 4u + base_code(kmer[i]);
    return r;
}
void dnbo_kmer_ranks(const char *seq, size_t len, uint32_t *out) {
    if (len < KLEN) return;
    for (size_t i = 0; i + KLEN <= len; i++) out[i] = dnbo_kmer2index(seq + i, KLEN);
}

/* ------------------------------------------------------------------------------------------------
 * a8  quantile scaling                      src/event_handling.cpp:451-541
 * ---------------------------------------------------------------------------------------------- */
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
/* event_handling.cpp:451-475 */
static void quantile_medians(const double *data, size_t n, double *q) {
    double *s = (double *)malloc(n * sizeof(double));
    memcpy(s, data, n * sizeof(double));
    qsort(s, n, sizeof(double), cmp_double);
    unsigned int m = (unsigned int)(n / 10);
    for (int i = 0; i < 10; i++) q[i] = s[((unsigned)i * m + (unsigned)(i + 1) * m) / 2];
    free(s);
}
int dnbo_quantile_scaling(const double *ev_mean, size_t n_events, const uint32_t *rank_ref, size_t n_ref,
                          const double *model_mean, double *shift, double *scale) {
    if (n_events == 0 || n_ref == 0) return 1;
    double *mm = (double *)malloc(n_ref * sizeof(double));
    for (size_t i = 0; i < n_ref; i++) mm[i] = model_mean[rank_ref[i]];
    double y[10], x[10];
    quantile_medians(ev_mean, n_events, y); /* signal quantiles */
    quantile_medians(mm, n_ref, x);         /* model quantiles */
    free(mm);
    /* linear_regression(x = model, y = signal), event_handling.cpp:478-507 */
    double sx = 0., sx2 = 0., sy = 0., sxy = 0.;
    int n = 10;
    for (int i = 0; i < n; i++) {
        sx = sx + x[i];
        sx2 = sx2 + x[i] * x[i];
        sy = sy + y[i];
        sxy = sxy + x[i] * y[i];
    }
    double slope = (n * sxy - sx * sy) / (n * sx2 - sx * sx);
    double icpt = (sy - slope * sx) / n;
    *shift = icpt;
    *scale = slope;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * a9  emission                              src/event_handling.cpp:116-137
 * ---------------------------------------------------------------------------------------------- */
static inline float emission(double x, double mu, double sigma, double ln_sigma) {
    float a = (float)((x - mu) / sigma);
    const float log_inv_sqrt_2pi = (float)log(0.3989422804014327);
    double lp = ((double)log_inv_sqrt_2pi - ln_sigma) + (double)((-0.5f * a) * a);
    return (float)lp;
}
float dnbo_log_probability_match(double ev_mean, double shift, double scale, double mu, double sigma) {
    double x = (ev_mean - shift) / scale;
    return emission(x, mu, sigma, log(sigma));
}

/* ------------------------------------------------------------------------------------------------
 * std::sort on doubles, restated (libstdc++, GCC 13, bits/stl_algo.h + bits/stl_heap.h): introsort (median-of-3
 * pivot moved to the front, unguarded Hoare partition, heapsort when 2*lg(n) levels are used up) and the final
 * insertion sort with its 16-element threshold.
 * Why a sort needs restating: a 0/0 Theil-Sen slope is a NaN, operator< is not a strict weak order with a NaN in
 * the range, and where the NaN ends up -- before or after element [size/2], i.e. which slope becomes the median --
 * is whatever this algorithm's sequence of swaps does (event_handling.cpp:77-78; about one read in 1000).  Without
 * a NaN every correct sort gives the same array, so this is the reference's behaviour in all cases.
 * ---------------------------------------------------------------------------------------------- */
static void ss_push_heap(double *first, long hole, long top, double value) {
    long parent = (hole - 1) / 2;
    while (hole > top && first[parent] < value) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void ss_adjust_heap(double *first, long hole, long len, double value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (first[child] < first[child - 1]) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    ss_push_heap(first, hole, top, value);
}
static void ss_heapsort(double *first, double *last) {          /* std::__partial_sort(first, last, last) */
    const long len = last - first;
    if (len >= 2)
        for (long parent = (len - 2) / 2;; parent--) {           /* std::__make_heap */
            ss_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
        }
    while (last - first > 1) {                                   /* std::__sort_heap / __pop_heap */
        --last;
        const double value = *last;
        *last = *first;
        ss_adjust_heap(first, 0, last - first, value);
    }
}
static void ss_swap(double *a, double *b) { const double t = *a; *a = *b; *b = t; }
static void ss_move_median_to_first(double *result, double *a, double *b, double *c) {
    if (*a < *b) {
        if (*b < *c) ss_swap(result, b);
        else if (*a < *c) ss_swap(result, c);
        else ss_swap(result, a);
    } else if (*a < *c) ss_swap(result, a);
    else if (*b < *c) ss_swap(result, c);
    else ss_swap(result, b);
}
static double *ss_unguarded_partition(double *first, double *last, const double *pivot) {
    for (;;) {
        while (*first < *pivot) ++first;
        --last;
        while (*pivot < *last) --last;
        if (!(first < last)) return first;
        ss_swap(first, last);
        ++first;
    }
}
static void ss_introsort_loop(double *first, double *last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) { ss_heapsort(first, last); return; }
        --depth_limit;
        double *mid = first + (last - first) / 2;
        ss_move_median_to_first(first, first + 1, mid, last - 1);
        double *cut = ss_unguarded_partition(first + 1, last, first);
        ss_introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static void ss_unguarded_linear_insert(double *last) {
    const double val = *last;
    double *next = last - 1;
    while (val < *next) { *last = *next; last = next; --next; }
    *last = val;
}
static void ss_insertion_sort(double *first, double *last) {
    if (first == last) return;
    for (double *i = first + 1; i != last; ++i) {
        if (*i < *first) {
            const double val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(double));
            *first = val;
        } else ss_unguarded_linear_insert(i);
    }
}
static void stdsort_double(double *first, size_t n) {
    if (n == 0) return;
    double *last = first + n;
    long lg = 0;
    for (size_t k = n; k > 1; k >>= 1) lg++;
    ss_introsort_loop(first, last, 2 * lg);
    if (n > 16) {
        ss_insertion_sort(first, first + 16);
        for (double *i = first + 16; i != last; ++i) ss_unguarded_linear_insert(i);
    } else ss_insertion_sort(first, last);
}

/* ------------------------------------------------------------------------------------------------
 * a12 Theil-Sen refinement                  src/event_handling.cpp:24-110
 * ---------------------------------------------------------------------------------------------- */
void dnbo_theil_sen(const double *sig, const uint32_t *ranks, size_t n, const double *model_mean, double shift,
                    double scale, double *out_shift, double *out_scale) {
    const size_t maxPoints = 1000, trim = 50;
    *out_shift = shift;
    *out_scale = scale;
    if (n < maxPoints) return;                          /* :33 */
    size_t eff = n - 2 * trim, skip = 1, np = eff;
    if (eff > maxPoints) {
        skip = eff / maxPoints;
        np = maxPoints;
    }
    double *x = (double *)malloc(np * sizeof(double)), *y = (double *)malloc(np * sizeof(double));
    size_t i = trim;
    for (size_t j = 0; j < np; j++) {
        x[j] = (sig[i] - shift) / scale;
        y[j] = model_mean[ranks[i]];
        i += skip;
    }
    size_t ns = np * (np - 1) / 2, c = 0;
    double *sl = (double *)malloc((ns ? ns : 1) * sizeof(double));
    for (size_t a = 0; a < np; a++)
        for (size_t b = a + 1; b < np; b++) sl[c++] = (y[a] - y[b]) / (x[a] - x[b]);
    stdsort_double(sl, ns);                              /* :77 std::sort, NaN behaviour included */
    double slope = sl[ns / 2];
    double *ic = (double *)malloc(np * sizeof(double));
    for (size_t a = 0; a < np; a++) ic[a] = y[a] - slope * x[a];
    stdsort_double(ic, np);                              /* :86 */
    double icpt = ic[np / 2];
    if (slope == 0.) {
        *out_shift = -1.;
        *out_scale = -1.;
    } else {
        double scale_corr = 1. / slope;
        double shift_corr = -icpt / slope;
        *out_shift = shift + (shift_corr * scale);
        *out_scale = scale * scale_corr;
    }
    free(x); free(y); free(sl); free(ic);
}

/* ------------------------------------------------------------------------------------------------
 * a10-a11 adaptive banded alignment         src/event_handling.cpp:148-448
 *
 * Representation (differs from the reference on purpose; verified identical): three
 * rolling 100-float bands, x_e per event, mu_k per k-mer, one move bit per band, one trace code per cell,
 * and the score of the last-k-mer column per event.  The lower-left (event,kmer) of every band is kept.
 * ---------------------------------------------------------------------------------------------- */
enum { FROM_D = 0, FROM_U = 1, FROM_L = 2 };

static int banded_align(const double *ev_mean, size_t E, const uint32_t *rq, size_t K, const uint32_t *rr, size_t Kref,
                        const int32_t *q2r, const double *model_mean, const double *model_stdv, double shift,
                        double scale, int keep_bands, dnbo_result *o) {
    const int bw = W, half = W / 2;
    size_t n_bands = (E + 1) + (K + 1);
    /* transition penalties :174-183 */
    double epk = (double)E / (double)K;
    double p_stay = 1 - (1 / (epk + 1));
    double lp_skip = log(1e-30), lp_stay = log(p_stay);
    double lp_step = log(1.0 - exp(lp_skip) - exp(lp_stay)), lp_trim = log(0.01);
    o->lp_skip = lp_skip; o->lp_stay = lp_stay; o->lp_step = lp_step; o->lp_trim = lp_trim;
    o->n_bands = n_bands;

    double *x = (double *)malloc(E * sizeof(double));
    for (size_t e = 0; e < E; e++) x[e] = (ev_mean[e] - shift) / scale; /* event_handling.cpp:130 */
    double *mu = (double *)malloc(K * sizeof(double)), *sg = (double *)malloc(K * sizeof(double)),
           *lsg = (double *)malloc(K * sizeof(double));
    double ln014 = log(0.14);
    for (size_t k = 0; k < K; k++) {
        mu[k] = model_mean[rq[k]];
        sg[k] = model_stdv ? model_stdv[rq[k]] : 0.14;
        lsg[k] = model_stdv ? log(sg[k]) : ln014;
    }
    int *ll_e = (int *)malloc(n_bands * sizeof(int)), *ll_k = (int *)malloc(n_bands * sizeof(int));
    uint8_t *trace = (uint8_t *)calloc(n_bands * W, 1);
    uint8_t *mv = (uint8_t *)calloc(n_bands, 1);
    float *lastcol = (float *)malloc((E ? E : 1) * sizeof(float));
    for (size_t e = 0; e < E; e++) lastcol[e] = -INFINITY;
    float bandbuf[3][W];
    float *b2 = bandbuf[0], *b1 = bandbuf[1], *b0 = bandbuf[2];
    for (int i = 0; i < W; i++) b2[i] = b1[i] = -INFINITY;
    /* :213-228 */
    ll_e[0] = half - 1; ll_k[0] = -1 - half;
    ll_e[1] = ll_e[0] + 1; ll_k[1] = ll_k[0];
    b2[(-1) - ll_k[0]] = 0.0f;
    {
        int off = ll_e[1] - 0;
        b1[off] = (float)lp_trim;
        trace[1 * W + off] = FROM_U;
    }
    int64_t fills = 0;
    for (size_t b = 2; b < n_bands; b++) {
        float ll = b1[0], ur = b1[bw - 1];
        int ll_ob = ll == -INFINITY, ur_ob = ur == -INFINITY;
        int right = (ll_ob && ur_ob) ? (b % 2 == 1) : (ll < ur); /* :243-247 */
        mv[b] = (uint8_t)right;
        if (right) { ll_e[b] = ll_e[b - 1]; ll_k[b] = ll_k[b - 1] + 1; }
        else       { ll_e[b] = ll_e[b - 1] + 1; ll_k[b] = ll_k[b - 1]; }
        for (int i = 0; i < W; i++) b0[i] = -INFINITY;
        int trim_off = (-1) - ll_k[b];                     /* :256-265 */
        if (trim_off >= 0 && trim_off < bw) {
            long ev = (long)ll_e[b] - trim_off;            /* unsigned wrap in the reference == "negative is out" */
            if (ev >= 0 && (size_t)ev < E) {
                b0[trim_off] = (float)(lp_trim * (double)((unsigned)ev + 1u));
                trace[b * W + trim_off] = FROM_U;
            }
        }
        int kmin = 0 - ll_k[b], kmax = (int)K - ll_k[b];
        int emin = ll_e[b] - ((int)E - 1), emax = ll_e[b] - (-1);
        int lo = kmin > emin ? kmin : emin; if (lo < 0) lo = 0;
        int hi = kmax < emax ? kmax : emax; if (hi > bw) hi = bw;
        for (int off = lo; off < hi; off++) {
            int ev = ll_e[b] - off, km = ll_k[b] + off;
            int o_up = ll_e[b - 1] - (ev - 1);
            int o_left = (km - 1) - ll_k[b - 1];
            int o_diag = (km - 1) - ll_k[b - 2];
            float up = (o_up >= 0 && o_up < bw) ? b1[o_up] : -INFINITY;
            float left = (o_left >= 0 && o_left < bw) ? b1[o_left] : -INFINITY;
            float diag = (o_diag >= 0 && o_diag < bw) ? b2[o_diag] : -INFINITY;
            float em = emission(x[ev], mu[km], sg[km], lsg[km]);
            float sd = (float)(((double)diag + lp_step) + (double)em);   /* :296-298 */
            float su = (float)(((double)up + lp_stay) + (double)em);
            float sl = (float)((double)left + lp_skip);
            float m = sd; uint8_t from = FROM_D;                          /* :300-306 */
            m = su > m ? su : m; from = (m == su) ? FROM_U : from;
            m = sl > m ? sl : m; from = (m == sl) ? FROM_L : from;
            b0[off] = m;
            trace[b * W + off] = from;
            fills++;
            if (km == (int)K - 1) lastcol[ev] = m;
        }
        float *t = b2; b2 = b1; b1 = b0; b0 = t;
    }
    o->fills = fills;

    /* end cell :324-340 */
    float max_score = -INFINITY;
    int cur_e = 0, cur_k = (int)K - 1, found = 0;
    for (size_t e = 0; e < E; e++) {
        size_t b = (e + 1) + (size_t)(cur_k + 1);
        int off = ll_e[b] - (int)e;
        if (off >= 0 && off < bw) {
            float s = (float)((double)lastcol[e] + (double)(E - e) * lp_trim);
            if (s > max_score) { max_score = s; cur_e = (int)e; found = 1; }
        }
    }
    o->end_score = max_score;
    int rc = 0;
    size_t cap = n_bands + 2;
    o->align_event = (uint32_t *)malloc(cap * sizeof(uint32_t));
    o->align_kmer = (uint32_t *)malloc(cap * sizeof(uint32_t));
    o->cleaned_signal = (double *)malloc((K + 1) * sizeof(double));
    o->cleaned_rank = (uint32_t *)malloc((K + 1) * sizeof(uint32_t));
    size_t na = 0, nc = 0;
    if (!found) {
        /* the reference would read trace[] out of range here (undefined behaviour) */
        rc = DNBO_UNDEFINED;
    } else {
        /* backtrace :356-412 */
        double sum_em = 0., n_aligned = 0.;
        int gap = 0, max_gap = 0;
        double buf_total = 0.; size_t buf_n = 0; /* signalBuffer: only its running sum (push order) and size matter */
        while (cur_k >= 0 && cur_e >= 0) {
            o->align_event[na] = (uint32_t)cur_e; o->align_kmer[na] = (uint32_t)cur_k; na++;
            float lp = emission(x[cur_e], mu[cur_k], sg[cur_k], lsg[cur_k]);
            sum_em += lp;
            n_aligned += 1;
            size_t b = (size_t)(cur_e + 1) + (size_t)(cur_k + 1);
            int off = ll_e[b] - cur_e;
            if (off < 0 || off >= bw) { rc = DNBO_UNDEFINED; break; }
            uint8_t from = trace[b * W + off];
            if (from == FROM_D) {
                buf_total += ev_mean[cur_e]; buf_n++;
                if (q2r[cur_k] >= 0) {                                 /* queryToRef.count() :386 */
                    uint32_t pr = (uint32_t)q2r[cur_k];
                    if (pr < Kref) {
                        o->cleaned_rank[nc] = rr[pr];
                        o->cleaned_signal[nc] = buf_total / (double)buf_n; /* vectorMean, common.h:184-194 */
                        nc++;
                    }
                }
                buf_total = 0.; buf_n = 0;
                cur_k--; cur_e--; gap = 0;
            } else if (from == FROM_U) {
                buf_total += ev_mean[cur_e]; buf_n++;
                cur_e--; gap = 0;
            } else {
                cur_k--; gap++;
                if (gap > max_gap) max_gap = gap;
            }
        }
        /* reverse :413 */
        for (size_t i = 0; i < na / 2; i++) {
            uint32_t t = o->align_event[i]; o->align_event[i] = o->align_event[na - 1 - i]; o->align_event[na - 1 - i] = t;
            t = o->align_kmer[i]; o->align_kmer[i] = o->align_kmer[na - 1 - i]; o->align_kmer[na - 1 - i] = t;
        }
        o->avg_log_emission = sum_em / n_aligned;
        o->spanned = na > 0 && o->align_kmer[0] == 0 && o->align_kmer[na - 1] == (uint32_t)(K - 1);
        o->max_gap = max_gap;
        if (rc == 0) {
            if (o->avg_log_emission < -2.0 || !o->spanned || max_gap > 5) rc = DNBO_QC_FAIL;   /* :433 */
            else if (nc < 1000) rc = DNBO_QC_FAIL;                                             /* :438 */
        }
    }
    o->n_align = na;
    o->n_cleaned = nc;
    if (keep_bands) {
        o->band_move = mv; o->trace = trace; o->last_col = lastcol;
    } else {
        free(mv); free(trace); free(lastcol);
    }
    free(x); free(mu); free(sg); free(lsg); free(ll_e); free(ll_k);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * a13  normaliseEvents                      src/event_handling.cpp:544-607
 * ---------------------------------------------------------------------------------------------- */
int dnbo_normalise(const float *raw, size_t n_raw, const char *query, size_t qlen, const char *ref, size_t rlen,
                   const int32_t *query_to_ref, const double *model_mean, const double *model_stdv, int keep_bands,
                   dnbo_result *o) {
    memset(o, 0, sizeof(*o));
    o->status = DNBO_UNDEFINED;
    if (qlen <= KLEN || rlen < KLEN || n_raw == 0) return o->status;
    dnbo_detector_param p = {3, 6, 1.4f, 9.0f, 0.2f};
    size_t cap = n_raw + 2;
    o->et_start = (uint64_t *)malloc(cap * sizeof(uint64_t));
    o->et_length = (float *)malloc(cap * sizeof(float));
    o->et_mean = (float *)malloc(cap * sizeof(float));
    o->et_stdv = (float *)malloc(cap * sizeof(float));
    o->et_n = dnbo_detect_events(raw, n_raw, p, o->et_start, o->et_length, o->et_mean, o->et_stdv, cap);
    if (o->et_n == 0) return o->status;
    /* :549-575 (quirks Q1-Q3) */
    o->ev_mean = (double *)malloc(o->et_n * sizeof(double));
    o->ev_start = (uint32_t *)malloc((o->et_n + 1) * sizeof(uint32_t));
    size_t E = 0;
    double mean = 0.;
    uint32_t raw_start = 0;
    for (size_t i = 0; i < o->et_n; i++) {
        if ((double)o->et_mean[i] > 0. && i > 0) {
            o->ev_mean[E] = mean;
            o->ev_start[E] = raw_start;
            E++;
            mean = (double)o->et_mean[i];
            raw_start = (uint32_t)o->et_start[i];
        }
    }
    o->ev_start[E] = raw_start;
    o->n_events = E;
    size_t K = qlen - KLEN + 1, Kref = rlen - KLEN + 1;
    o->n_kmers = K; o->n_kmers_ref = Kref;
    o->rank_query = (uint32_t *)malloc(K * sizeof(uint32_t));
    o->rank_ref = (uint32_t *)malloc(Kref * sizeof(uint32_t));
    dnbo_kmer_ranks(query, qlen, o->rank_query);
    dnbo_kmer_ranks(ref, rlen, o->rank_ref);
    o->events_per_base = (double)o->et_n / (double)(qlen - KLEN);   /* :606 (Q4) */
    if (E == 0) return o->status;
    if (dnbo_quantile_scaling(o->ev_mean, E, o->rank_ref, Kref, model_mean, &o->rough_shift, &o->rough_scale))
        return o->status;
    o->shift = o->rough_shift; o->scale = o->rough_scale;
    int rc = banded_align(o->ev_mean, E, o->rank_query, K, o->rank_ref, Kref, query_to_ref, model_mean, model_stdv,
                          o->rough_shift, o->rough_scale, keep_bands, o);
    if (rc == DNBO_UNDEFINED) return o->status;
    dnbo_theil_sen(o->cleaned_signal, o->cleaned_rank, o->n_cleaned, model_mean, o->rough_shift, o->rough_scale,
                   &o->shift, &o->scale);
    if (o->shift == -1.) rc = rc ? rc : DNBO_SCALE_FAIL;             /* :604 */
    o->status = rc;
    return rc;
}

void dnbo_result_free(dnbo_result *r) {
    free(r->et_start); free(r->et_length); free(r->et_mean); free(r->et_stdv);
    free(r->ev_mean); free(r->ev_start); free(r->rank_query); free(r->rank_ref);
    free(r->band_move); free(r->trace); free(r->last_col);
    free(r->align_event); free(r->align_kmer); free(r->cleaned_signal); free(r->cleaned_rank);
    memset(r, 0, sizeof(*r));
}

double dnbo_bench_normalise(const float *const *raw, const size_t *n_raw, const char *const *query, const size_t *qlen,
                            const char *const *ref, const size_t *rlen, const int32_t *const *q2r, size_t n_reads,
                            const double *model_mean, int threads, int *failed) {
    int nfail = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(+ : nfail)
    for (size_t i = 0; i < n_reads; i++) {
        dnbo_result r;
        int rc = dnbo_normalise(raw[i], n_raw[i], query[i], qlen[i], ref[i], rlen[i], q2r[i], model_mean, NULL, 0, &r);
        if (rc != DNBO_OK) nfail++;
        dnbo_result_free(&r);
    }
    double t1 = omp_get_wtime();
    if (failed) *failed = nfail;
    return t1 - t0;
}

/* ------------------------------------------------------------------------------------------------
 * a14  log-space helpers                    src/probability.cpp:23-154   (NaN == log 0)
 * ---------------------------------------------------------------------------------------------- */
double dnbo_eexp(double x) { return isnan(x) ? 0.0 : exp(x); }
int dnbo_eln(double x, double *out) {
    if (x == 0.0) { *out = NAN; return 0; }
    if (x > 0.0) { *out = log(x); return 0; }
    return 1; /* NegativeLog (also for NaN input, which fails both tests) */
}
static double eln_(double x) { double o = NAN; dnbo_eln(x, &o); return o; }
double dnbo_lnSum(double a, double b) {
    if (isnan(a) || isnan(b)) {
        if (isnan(a) && isnan(b)) return NAN;
        return isnan(a) ? b : a;
    }
    if (a > b) return a + eln_(1.0 + dnbo_eexp(b - a));
    return b + eln_(1.0 + dnbo_eexp(a - b));
}
double dnbo_lnProd(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : a + b; }
int dnbo_lnGreaterThan(double a, double b) {
    if (isnan(a) || isnan(b)) {
        if (isnan(a) || !isnan(b)) return 0;   /* probability.cpp:112 */
        return 1;                              /* :115 (a finite, b NaN) */
    }
    return a > b;
}
double dnbo_uniformPDF(double lb, double ub, double x) { return (x >= lb && x <= ub) ? 1.0 / (ub - lb) : 0.0; }
double dnbo_normalPDF(double mu, double sigma, double x) {
    return (1.0 / sqrt(2.0 * pow(sigma, 2.0) * M_PI)) * exp(-pow(x - mu, 2.0) / (2.0 * pow(sigma, 2.0)));
}
double dnbo_cauchyPDF(double loc, double scale, double x) {
    return 1. / ((scale * M_PI) * (1. + pow((x - loc) / scale, 2.)));
}

/* ------------------------------------------------------------------------------------------------
 * a15  analogue forward likelihood          src/detect.cpp:235-378
 * ---------------------------------------------------------------------------------------------- */
#define LS dnbo_lnSum
#define LP dnbo_lnProd
double dnbo_sequence_probability(const double *obs, size_t n_obs, const char *seq, size_t seq_len, size_t window,
                                 int use_analogue, double shift, double scale, double events_per_base,
                                 size_t a_start, size_t a_end, const double *unl_mean, const double *unl_stdv,
                                 const double *ana_mean, const double *ana_stdv) {
    (void)seq_len;
    const size_t n = 2 * window;
    /* HMM_TransitionProbs_DNA_R10 = {0.3, 0.7, 0.999, 0.0025, 0.001, 0.001}, config.h:42 */
    double externalD2D = eln_(0.3), externalD2M1 = eln_(0.7), externalI2M1 = eln_(0.999), externalM12D = eln_(0.0025);
    double internalM12I = eln_(0.001), internalI2I = eln_(0.001);
    double internalM12M1 = eln_(1. - (1. / events_per_base));
    double externalM12M1 = eln_(1.0 - externalM12D - internalM12I - internalM12M1); /* Q10: logs, on purpose */
    double *buf = (double *)malloc(6 * n * sizeof(double));
    double *Ic = buf, *Dc = buf + n, *Mc = buf + 2 * n, *Ip = buf + 3 * n, *Dp = buf + 4 * n, *Mp = buf + 5 * n;
    for (size_t i = 0; i < 6 * n; i++) buf[i] = NAN;
    double firstI_curr = NAN, firstI_prev = NAN, start_curr = NAN, start_prev = 0.0;
    Dp[0] = LP(start_prev, eln_(0.25));
    for (size_t i = 1; i < n; i++) Dp[i] = LP(Dp[i - 1], externalD2D);
    for (size_t t = 0; t < n_obs; t++) {
        for (size_t i = 0; i < n; i++) Ic[i] = Mc[i] = Dc[i] = NAN;
        firstI_curr = NAN;
        double xo = (obs[t] - shift) / scale;
        uint32_t r0 = dnbo_kmer2index(seq, KLEN);
        double matchProb = eln_(dnbo_normalPDF(unl_mean[r0], unl_stdv[r0], xo));
        double insProb = 0.0;
        firstI_curr = LS(firstI_curr, LP(LP(start_prev, eln_(0.25)), insProb));
        firstI_curr = LS(firstI_curr, LP(LP(firstI_prev, eln_(0.25)), insProb));
        Ic[0] = LS(Ic[0], LP(LP(Ip[0], internalI2I), insProb));
        Ic[0] = LS(Ic[0], LP(LP(Mp[0], internalM12I), insProb));
        Mc[0] = LS(Mc[0], LP(LP(firstI_prev, eln_(0.5)), matchProb));
        Mc[0] = LS(Mc[0], LP(LP(Mp[0], internalM12M1), matchProb));
        Mc[0] = LS(Mc[0], LP(LP(start_prev, eln_(0.5)), matchProb));
        Dc[0] = LS(Dc[0], LP(NAN, eln_(0.25)));
        Dc[0] = LS(Dc[0], LP(firstI_curr, eln_(0.25)));
        for (size_t i = 1; i < n; i++) {
            const char *km = seq + i;
            uint32_t rk = dnbo_kmer2index(km, KLEN);
            int hasT = memchr(km, 'T', KLEN) != NULL;
            if (use_analogue && a_start <= i && i <= a_end && hasT)
                matchProb = eln_(dnbo_normalPDF(ana_mean[rk], ana_stdv[rk], xo));
            else
                matchProb = eln_(dnbo_normalPDF(unl_mean[rk], unl_stdv[rk], xo));
            Ic[i] = LS(Ic[i], LP(LP(Ip[i], internalI2I), insProb));
            Ic[i] = LS(Ic[i], LP(LP(Mp[i], internalM12I), insProb));
            Mc[i] = LS(Mc[i], LP(LP(Ip[i - 1], externalI2M1), matchProb));
            Mc[i] = LS(Mc[i], LP(LP(Mp[i - 1], externalM12M1), matchProb));
            Mc[i] = LS(Mc[i], LP(LP(Mp[i], internalM12M1), matchProb));
            Mc[i] = LS(Mc[i], LP(LP(Dp[i - 1], externalD2M1), matchProb));
        }
        for (size_t i = 1; i < n; i++) {
            Dc[i] = LS(Dc[i], LP(Mc[i - 1], externalM12D));
            Dc[i] = LS(Dc[i], LP(Dc[i - 1], externalD2D));
        }
        memcpy(Ip, Ic, n * sizeof(double));
        memcpy(Mp, Mc, n * sizeof(double));
        memcpy(Dp, Dc, n * sizeof(double));
        firstI_prev = firstI_curr;
        start_prev = start_curr;
    }
    double fwd = NAN;
    fwd = LS(fwd, LP(Dc[n - 1], eln_(1.0)));
    fwd = LS(fwd, LP(Mc[n - 1], LS(externalM12M1, externalM12D)));
    fwd = LS(fwd, LP(Ic[n - 1], externalI2M1));
    free(buf);
    return fwd;
}

/* ------------------------------------------------------------------------------------------------------------
 * SURVEY s.8 row f1: builtinViterbi (src/alignment.cpp:193-516) and eventalign (src/alignment.cpp:547-744)
 * ------------------------------------------------------------------------------------------------------------ */
static int gt_(double a, double b) { return dnbo_lnGreaterThan(a, b); }

/* alignment.cpp:193-516 with flip == false.  Backtrace codes instead of the two size_t matrices:
 *   I: 0 = from I(i)   1 = from M(i)   2 = from start (i == 0 only)
 *   M: 0 = from I(i-1) 1 = from M(i-1) 2 = from M(i)  3 = from D(i-1);  for i == 0: 0 = from M(0), 1 = from start
 *   D: 0 = from M(i-1) 1 = from D(i-1);  D(0) always comes from start */
size_t dnbo_builtin_viterbi(const double *obs, size_t T, const char *seq, size_t seq_len, double shift, double scale,
                            double events_per_base, const double *model_mean, const double *model_stdv, double *score,
                            int32_t *idx, uint8_t *type, size_t cap) {
    const double externalD2D = eln_(0.3), externalD2M1 = eln_(0.7), externalI2M1 = eln_(0.999), externalM12D = eln_(0.0025);
    const double internalM12I = eln_(0.001), internalI2I = eln_(0.001);
    const double internalM12M1 = eln_(1. - (1. / events_per_base));                                   /* :208 */
    const double externalM12M1 = eln_(1.0 - externalM12D - internalM12I - internalM12M1);             /* :209 (Q10) */
    const double externalM12M1orD = dnbo_lnSum(externalM12M1, externalM12D);
    const double externalOrInternalM12M1 = dnbo_lnSum(externalM12M1, internalM12M1);
    const size_t n = seq_len - KLEN + 1;
    double *buf = (double *)malloc(8 * n * sizeof(double));
    double *Ip = buf, *Mp = buf + n, *Dp = buf + 2 * n, *Ic = buf + 3 * n, *Mc = buf + 4 * n, *Dc = buf + 5 * n;
    double *mu = buf + 6 * n, *sg = buf + 7 * n;
    uint8_t *bI = (uint8_t *)calloc(3 * n * (T + 1), 1), *bM = bI + n * (T + 1), *bD = bM + n * (T + 1);
    for (size_t i = 0; i < n; i++) {
        uint32_t rk = dnbo_kmer2index(seq + i, KLEN);
        mu[i] = model_mean[rk];
        sg[i] = model_stdv ? model_stdv[rk] : 0.14;
        Ip[i] = Mp[i] = Dp[i] = Ic[i] = Mc[i] = Dc[i] = NAN;
    }
    double start_prev = 0.0;
    const double start_curr = NAN;
    Dp[0] = dnbo_lnProd(start_prev, externalM12D);                                                    /* :241 */
    for (size_t i = 1; i < n; i++) Dp[i] = Dp[i - 1] + externalD2D;                                   /* :248 */
    for (size_t t = 0; t < T; t++) {
        for (size_t i = 0; i < n; i++) Ic[i] = Mc[i] = Dc[i] = NAN;
        const double x = (obs[t] - shift) / scale;
        uint8_t *cI = bI + (t + 1) * n, *cM = bM + (t + 1) * n, *cD = bD + (t + 1) * n;
        double mp = eln_(dnbo_normalPDF(mu[0], sg[0], x));
        {   /* base 1 insertion :276-300 */
            const double v0 = Ip[0] + internalI2I + 0.0, v1 = Mp[0] + internalM12I + 0.0, v2 = start_prev + internalM12I + 0.0;
            double m = v0; int a = 0;
            if (gt_(v1, m)) { m = v1; a = 1; }
            if (gt_(v2, m)) { m = v2; a = 2; }
            Ic[0] = m; cI[0] = (uint8_t)a;
        }
        {   /* base 1 match :303-322 */
            const double v0 = Mp[0] + internalM12M1 + mp, v1 = start_prev + externalOrInternalM12M1 + mp;
            double m = v0; int a = 0;
            if (gt_(v1, m)) { m = v1; a = 1; }
            Mc[0] = m; cM[0] = (uint8_t)a;
        }
        Dc[0] = dnbo_lnProd(NAN, externalM12D);                                                       /* :325 */
        for (size_t i = 1; i < n; i++) {
            mp = eln_(dnbo_normalPDF(mu[i], sg[i], x));
            {
                const double v0 = Ip[i] + internalI2I + 0.0, v1 = Mp[i] + internalM12I + 0.0;         /* :350-356 */
                double m = v0; int a = 0;
                if (gt_(v1, m)) { m = v1; a = 1; }
                Ic[i] = m; cI[i] = (uint8_t)a;
            }
            {
                const double v0 = Ip[i - 1] + externalI2M1 + mp, v1 = Mp[i - 1] + externalM12M1 + mp;  /* :372-381 */
                const double v2 = Mp[i] + internalM12M1 + mp, v3 = Dp[i - 1] + externalD2M1 + mp;
                double m = v0; int a = 0;
                if (gt_(v1, m)) { m = v1; a = 1; }
                if (gt_(v2, m)) { m = v2; a = 2; }
                if (gt_(v3, m)) { m = v3; a = 3; }
                Mc[i] = m; cM[i] = (uint8_t)a;
            }
        }
        for (size_t i = 1; i < n; i++) {                                                              /* :405-428 */
            const double v0 = Mc[i - 1] + externalM12D, v1 = Dc[i - 1] + externalD2D;
            double m = v0; int a = 0;
            if (gt_(v1, m)) { m = v1; a = 1; }
            Dc[i] = m; cD[i] = (uint8_t)a;
        }
        memcpy(Ip, Ic, n * sizeof(double)); memcpy(Mp, Mc, n * sizeof(double)); memcpy(Dp, Dc, n * sizeof(double));
        start_prev = start_curr;
    }
    /* termination :447-474.  T == 0 leaves the *_curr vectors NaN: D wins by default (index 0) */
    double v0 = Dc[n - 1], v1 = Mc[n - 1] + externalM12M1orD, v2 = Ic[n - 1] + externalI2M1;
    double m = v0; int a = 0;
    if (gt_(v1, m)) { m = v1; a = 1; }
    if (gt_(v2, m)) { m = v2; a = 2; }
    *score = m;
    /* traceback :476-505; state space: type (0 D, 1 M, 2 I), index i, time t */
    int ty = a == 0 ? 0 : a == 1 ? 1 : 2;
    long i = (long)n - 1, t = (long)T;
    size_t np = 0;
    while (1) {
        if (np < cap) { idx[np] = (int32_t)i; type[np] = (uint8_t)ty; }
        np++;
        int nty; long ni = i, nt = t;
        if (ty == 0) {                       /* D: backtraceT = t (same column) */
            if (t == 0) { if (i == 0) break; nty = 0; ni = i - 1; }                     /* :242-250 */
            else if (i == 0) break;                                                     /* :326 */
            else { const int c = bD[t * n + i]; nty = c == 0 ? 1 : 0; ni = i - 1; }
        } else if (ty == 1) {                /* M: consumed observation t-1 */
            /* column 0 of an M or I row is never written: the zero-initialised matrices say "D state 0, t = 0" */
            if (t == 0) { ty = 0; i = 0; continue; }
            const int c = bM[t * n + i]; nt = t - 1;
            if (i == 0) { if (c == 1) break; nty = 1; }
            else if (c == 0) { nty = 2; ni = i - 1; } else if (c == 1) { nty = 1; ni = i - 1; }
            else if (c == 2) { nty = 1; } else { nty = 0; ni = i - 1; }
        } else {
            if (t == 0) { ty = 0; i = 0; continue; }
            const int c = bI[t * n + i]; nt = t - 1;
            if (c == 0) nty = 2; else if (c == 1) nty = 1; else break;
        }
        ty = nty; i = ni; t = nt;
    }
    /* reverse into path order */
    const size_t nn = np < cap ? np : cap;
    for (size_t q = 0; q < nn / 2; q++) {
        int32_t ti = idx[q]; idx[q] = idx[nn - 1 - q]; idx[nn - 1 - q] = ti;
        uint8_t tt = type[q]; type[q] = type[nn - 1 - q]; type[nn - 1 - q] = tt;
    }
    free(buf); free(bI);
    return np;
}

static int defined_acgt(const char *s, size_t n) {                                                    /* :519-544 */
    for (size_t i = 0; i < n; i++)
        if (!(s[i] == 'A' || s[i] == 'T' || s[i] == 'G' || s[i] == 'C')) return 0;
    return 1;
}

/* alignment.cpp:547-744 up to (not including) the text formatting: one record per event the reference prints lines
 * for -- (index into r.events, position on referenceSeqMappedTo = reference_index + pos, label 1 = M / 2 = I, the
 * window's indelScore).  The printed coordinate is refStart + ref_pos + k/2 (fwd) or refEnd - ref_pos - k/2 - 1 (rev).
 * Returns the number of records (may exceed cap; only cap are stored). */
size_t dnbo_eventalign(const char *ref, size_t rlen, const int32_t *r2q, const uint32_t *al_e, const uint32_t *al_k,
                       size_t n_align, const double *ev_mean, double shift, double scale, double events_per_base,
                       unsigned total_window, const double *model_mean, uint32_t *rec_event, uint32_t *rec_refpos,
                       uint8_t *rec_label, int32_t *rec_indel, size_t cap) {
    const unsigned k = KLEN;
    size_t nrec = 0;
    long read_head = 0;
    unsigned reference_index = 0;
    double *snip = (double *)malloc((n_align + 1) * sizeof(double));
    uint32_t *snip_ev = (uint32_t *)malloc((n_align + 1) * sizeof(uint32_t));
    size_t path_cap = 4 * (n_align + 4 * (size_t)total_window) + 64;
    int32_t *pidx = (int32_t *)malloc(path_cap * sizeof(int32_t));
    uint8_t *ptyp = (uint8_t *)malloc(path_cap);
    /* design aid: DNBO_EA_WINDOW_LOG=<file> appends "ref_index window_len n_obs last_m_ref last_m_ev" per window */
    FILE *wlog = getenv("DNBO_EA_WINDOW_LOG") ? fopen(getenv("DNBO_EA_WINDOW_LOG"), "a") : NULL;
    while (reference_index < rlen - k + 1) {
        const unsigned bases_to_end = (unsigned)rlen - reference_index;
        unsigned wl = bases_to_end < total_window ? bases_to_end : total_window;
        if (bases_to_end > 1.5 * total_window) {                                                      /* :565-595 */
            const size_t bl = (size_t)(1.5 * wl);
            const char *bs = ref + reference_index;
            if (!defined_acgt(bs, bl)) { reference_index += wl; continue; }
            for (unsigned i = wl; i < 1.5 * wl - k - 1; i++) {
                const double m = model_mean[dnbo_kmer2index(bs + i, k)];
                const double mb = model_mean[dnbo_kmer2index(bs + i - 1, k)];
                const double mf = model_mean[dnbo_kmer2index(bs + i + 1, k)];
                if (fabs(m - mf) > 0.75 && fabs(m - mb) > 0.75) { wl = i + k; break; }
            }
        }
        const char *rs = ref + reference_index;
        if (!defined_acgt(rs, wl)) { reference_index += wl; continue; }
        const uint32_t lo = (uint32_t)r2q[reference_index], hi = (uint32_t)r2q[reference_index + wl - k + 1];
        size_t ns = 0;
        int first = 1;
        for (size_t j = (size_t)read_head; j < n_align; j++) {                                        /* :611-632 */
            if (lo <= al_k[j] && al_k[j] < hi) {
                if (first) { read_head = (long)j; first = 0; }
                const double em = ev_mean[al_e[j]];
                if (0. < em && em < 250.) { snip[ns] = em; snip_ev[ns] = al_e[j]; ns++; }
            }
            if (al_k[j] >= hi) break;
        }
        const int indel = ((int)hi - (int)lo) - (int)(wl - k + 1);                                    /* :635-638 */
        if (ns < 2) { reference_index += wl; continue; }
        double score;
        const size_t np = dnbo_builtin_viterbi(snip, ns, rs, wl, shift, scale, events_per_base, model_mean, NULL, &score,
                                               pidx, ptyp, path_cap);
        size_t last_m_ev = 0, last_m_ref = 0, ev = 0;
        for (size_t i = 0; i < np; i++) {                                                             /* :661-672 */
            if (ptyp[i] == 1) { last_m_ev = ev; last_m_ref = (size_t)pidx[i]; }
            if (ptyp[i] != 0) ev++;
        }
        ev = 0;
        for (size_t i = 0; i < np; i++) {                                                             /* :676-736 */
            if (ptyp[i] == 0) continue;
            if (ptyp[i] == 1 || (ptyp[i] == 2 && ev < last_m_ev)) {
                if (nrec < cap) {
                    rec_event[nrec] = snip_ev[ev]; rec_refpos[nrec] = reference_index + (uint32_t)pidx[i];
                    rec_label[nrec] = ptyp[i]; rec_indel[nrec] = indel;
                }
                nrec++;
            }
            ev++;
        }
        if (wlog) fprintf(wlog, "%u %u %zu %zu %zu\n", reference_index, wl, ns, last_m_ref, last_m_ev);
        read_head += (long)last_m_ev + 1;                                                             /* :740-741 */
        reference_index += (unsigned)last_m_ref + 1;
    }
    if (wlog) fclose(wlog);
    free(snip); free(snip_ev); free(pidx); free(ptyp);
    return nrec;
}

/* ---- f2: DNN input tensors (reads.h:288-372, 147-172, 109-138; consumer detect.cpp:586-649) ----------------------
 * Literal restatement of what eventalign's r.addSignal calls (alignment.cpp:723) leave in refCoordToAP and of the
 * tensor builders that read it.  Input = the eventalign records (dnbo_eventalign) + the raw signal; positions are
 * kept in a coordinate-sorted table standing for the std::map. */
typedef struct {
    uint32_t coord, query_idx, ref_idx, ref_pos;
    int32_t quality;
    uint32_t n_sig;                 /* signals added (unbounded in the reference; only the first RAWDEPTH are read) */
    float sig[DNBO_RAWDEPTH];
} ap_t;

static int ap_cmp(const void *a, const void *b) {
    const uint32_t x = ((const ap_t *)a)->coord, y = ((const ap_t *)b)->coord;
    return x < y ? -1 : x > y;
}

static int u32_contains(const uint32_t *v, size_t n, uint32_t x) {      /* r.refCoordToCalls.count(event_coord) */
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t m = (lo + hi) / 2; if (v[m] < x) lo = m + 1; else hi = m; }
    return lo < n && v[lo] == x;
}

/* base2index of AlignedPosition (reads.h:84): A0 T1 G2 C3; an unknown base reads 0 through std::map::operator[] */
static unsigned ap_base(char c) { return c == 'T' ? 1u : c == 'G' ? 2u : c == 'C' ? 3u : 0u; }

size_t dnbo_dnn_features(const char *ref, size_t rlen, const int32_t *r2q, int is_reverse, uint32_t ref_start,
                         uint32_t ref_end, const uint32_t *rec_event, const uint32_t *rec_refpos,
                         const uint8_t *rec_label, const int32_t *rec_indel, size_t n_rec, const double *raw,
                         const uint32_t *event_start, double shift, double scale, const uint32_t *called,
                         size_t n_called, float *signal, float *core, float *residual, uint32_t *coords,
                         uint32_t *ref_index, uint32_t *query_index, int32_t *quality, size_t cap) {
    const unsigned k = KLEN;
    (void)rlen;
    ap_t *tab = (ap_t *)malloc((n_rec + 1) * sizeof(ap_t));
    size_t P = 0;
    for (size_t q = 0; q < n_rec; q++) {
        if (rec_label[q] != 1) continue;                                                   /* alignment.cpp:706 */
        const uint32_t pos = rec_refpos[q];
        const uint32_t coord = is_reverse ? ref_end - pos - k / 2 - 1 : ref_start + pos + k / 2;   /* :646-648, 690-697 */
        if (u32_contains(called, n_called, coord)) continue;                               /* :711 */
        const uint32_t iref = pos + k / 2;                                                 /* :701-702 */
        for (uint32_t j = event_start[rec_event[q]]; j < event_start[rec_event[q] + 1]; j++) {
            const double scaled = (raw[j] - shift) / scale;                                /* :709 */
            size_t a = P;                                                                  /* reads.h:288-300 */
            while (a > 0 && tab[a - 1].coord != coord) a--;     /* refCoordToAP.count(refPos): newest first (usually a hit) */
            if (a > 0) a--;
            else {
                a = P;
                tab[P].coord = coord; tab[P].query_idx = (uint32_t)r2q[iref]; tab[P].ref_idx = iref; tab[P].ref_pos = pos;
                tab[P].quality = rec_indel[q]; tab[P].n_sig = 0;
                P++;
            }
            if (tab[a].n_sig < DNBO_RAWDEPTH) tab[a].sig[tab[a].n_sig] = (float)scaled;    /* reads.h:154-159 */
            tab[a].n_sig++;
        }
    }
    qsort(tab, P, sizeof(ap_t), ap_cmp);                                                   /* std::map order */
    for (size_t o = 0; o < P && o < cap; o++) {
        const ap_t *a = &tab[is_reverse ? P - 1 - o : o];                                  /* rbegin for "rev", reads.h:320-327 */
        for (unsigned i = 0; i < DNBO_RAWDEPTH; i++) signal[o * DNBO_RAWDEPTH + i] = i < a->n_sig ? a->sig[i] : 0.f;   /* :162-168 */
        const char *km = ref + a->ref_pos;
        unsigned c = 0, r = 0;
        for (unsigned i = 2; i < 7; i++) c = c * 4u + ap_base(km[i]);                      /* getCoreIndex, reads.h:109-121 */
        r = ((ap_base(km[0]) * 4u + ap_base(km[1])) * 4u + ap_base(km[7])) * 4u + ap_base(km[8]);   /* getResidualIndex :122-135 */
        core[o] = (float)(c + 1u);
        residual[o] = (float)(r + 1u);
        coords[o] = a->coord; ref_index[o] = a->ref_idx; query_index[o] = a->query_idx; quality[o] = a->quality;
    }
    free(tab);
    return P;
}

/* detect.cpp:381-390 + 393-574 */
size_t dnbo_ll_across_read(const char *ref, size_t rlen, const int32_t *r2q, int is_reverse,
                           const uint32_t *al_e, const uint32_t *al_k, size_t n_align, const double *ev_mean,
                           double shift, double scale, double events_per_base, unsigned w, const double *unl_mean,
                           const double *unl_stdv, const double *ana_mean, const double *ana_stdv, uint32_t *pos,
                           double *llr, size_t cap) {
    const unsigned k = KLEN;
    size_t ncalls = 0;
    if (rlen < 4 * (size_t)w + 1 || n_align == 0) return 0;
    size_t npoi = 0;
    uint32_t *poi = (uint32_t *)malloc(rlen * sizeof(uint32_t));
    for (size_t i = 2 * w; i < rlen - 2 * w; i++)
        if (ref[i] == 'T') poi[npoi++] = (uint32_t)i;
    long read_head = 0;
    if (is_reverse) {
        read_head = (long)n_align - 1;
        for (size_t i = 0; i < npoi / 2; i++) { uint32_t t = poi[i]; poi[i] = poi[npoi - 1 - i]; poi[npoi - 1 - i] = t; }
    }
    double *snip = (double *)malloc((n_align + 1) * sizeof(double));
    for (size_t pi = 0; pi < npoi; pi++) {
        uint32_t p = poi[pi];
        if ((size_t)p - w + 2 * w + k > rlen) continue; /* substr would be short -> not all of 2w+k bases */
        const char *sn = ref + p - w;
        int ok = 1;
        for (unsigned i = 0; i < 2 * w + k; i++) {
            char c = sn[i];
            if (!(c == 'A' || c == 'T' || c == 'G' || c == 'C')) ok = 0;
        }
        if (!ok) continue;
        uint32_t lo = (uint32_t)r2q[p - w], hi = (uint32_t)r2q[p + w];
        size_t ns = 0;
        int first = 1;
        if (is_reverse) {
            for (long j = read_head; j >= 0; j--) {
                if (lo <= al_k[j] && al_k[j] < hi) {
                    if (first) { read_head = j; first = 0; }
                    double ev = ev_mean[al_e[j]];
                    if (ev > 0. && ev < 250.0) snip[ns++] = ev;
                }
                if (al_k[j] < lo) {
                    for (size_t i = 0; i < ns / 2; i++) { double t = snip[i]; snip[i] = snip[ns - 1 - i]; snip[ns - 1 - i] = t; }
                    break;
                }
            }
        } else {
            for (size_t j = (size_t)read_head; j < n_align; j++) {
                if (lo <= al_k[j] && al_k[j] < hi) {
                    if (first) { read_head = (long)j; first = 0; }
                    double ev = ev_mean[al_e[j]];
                    if (ev > 0. && ev < 250.0) snip[ns++] = ev;
                }
                if (al_k[j] >= hi) break;
            }
        }
        if (ns < 2 * w - k) continue;
        double la = dnbo_sequence_probability(snip, ns, sn, 2 * w + k, w, 1, shift, scale, events_per_base, w - k / 2,
                                              w + k / 2, unl_mean, unl_stdv, ana_mean, ana_stdv);
        double lt = dnbo_sequence_probability(snip, ns, sn, 2 * w + k, w, 0, shift, scale, events_per_base, 0, 0,
                                              unl_mean, unl_stdv, ana_mean, ana_stdv);
        if (ncalls < cap) { pos[ncalls] = p; llr[ncalls] = la - lt; }
        ncalls++;
    }
    free(poi); free(snip);
    return ncalls;
}
