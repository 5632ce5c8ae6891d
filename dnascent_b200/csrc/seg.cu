// seg.cu -- scrappie t-statistic event segmentation + the r.events filter, fused, one lane per read.
//
// Replaces (reference, paths relative to /root/reference):
//   compute_sum_sumsq            src/scrappie/event_detection.c:35-48
//   compute_tstat (w=3, w=6)     src/scrappie/event_detection.c:60-115
//   short_long_peak_detector     src/scrappie/event_detection.c:122-198
//   create_event(s)              src/scrappie/event_detection.c:213-266
//   event table -> r.events      src/event_handling.cpp:549-575   (quirks Q1-Q3)
//
// Why one lane per read: the double prefix sum of squares rounds at EVERY step (x*x needs up to 48 bits, the
// running sum passes 2^19 within ~100 samples), so its value depends on the serial order -- the survey
// measured 10 % of reads changing boundaries under a different summation.  Each lane therefore carries the
// reference's own serial chain; everything downstream of the sums (both t-statistics, both peak detectors,
// event construction and the r.events filter) is streamed in the same pass from a 16-deep ring of
// (sum, sumsq) kept in shared memory, so the signal is read from HBM exactly once and only the event table
// (u32 start + f32 mean per event) is written: 4 B/sample + 8 B/event algorithmic traffic.
#include <cfloat>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define SEG_THREADS 64
#define SEG_RING 16

namespace {

struct Detector {
    int pos;          // peak_pos, -1 = DEF_PEAK_POS
    float val;        // peak_value
    bool valid;       // valid_peak
    double psum, psq; // sums[peak_pos], sumsqs[peak_pos] (carried so that create_event needs no array)
};

struct EventSink {
    // scrappie event under construction starts at prev_peak
    uint32_t prev_peak;
    double prev_sum, prev_sq;
    uint32_t et_count;
    // event_handling.cpp:549-575 filter state
    uint32_t E;
    float pend_mean;
    uint32_t pend_start;
    // destinations
    uint32_t cap;
    uint32_t *ev_start;
    float *ev_mean;
    uint64_t *et_start;
    float *et_length, *et_mean, *et_stdv;
};

// event_detection.c:213-232 + event_handling.cpp:553-573 for one finished scrappie event [prev_peak, end)
__device__ __forceinline__ void finish_event(EventSink &k, uint32_t end, double esum, double esq) {
    unsigned long long span = (unsigned long long)((long long)end - (long long)k.prev_peak);  // size_t arithmetic
    float length = (float)span;
    float mean = fDiv(d2f(dSub(esum, k.prev_sum)), length);
    uint32_t idx = k.et_count++;
    if (k.et_start && idx <= k.cap) {
        float dsq = d2f(dSub(esq, k.prev_sq));
        float var = fSub(fDiv(dsq, length), fMul(mean, mean));
        k.et_start[idx] = k.prev_peak;
        k.et_length[idx] = length;
        k.et_mean[idx] = mean;
        k.et_stdv[idx] = __fsqrt_rn(fmaxf(var, 0.0f));
    }
    if ((double)mean > 0.0 && idx > 0) {
        if (k.E < k.cap) {
            k.ev_mean[k.E] = k.pend_mean;
            k.ev_start[k.E] = k.pend_start;
        }
        k.E++;
        k.pend_mean = mean;
        k.pend_start = k.prev_peak;
    }
    k.prev_peak = end;
    k.prev_sum = esum;
    k.prev_sq = esq;
}

// event_detection.c:89-112 for one position; sm/qm = sums at i-w, s0/q0 at i, sp/qp at i+w
__device__ __forceinline__ float tstat_at(double sm, double qm, double s0, double q0, double sp, double qp, float wf) {
    double sum1 = dSub(s0, sm);   // i == w subtracts sum[0] == 0, identical to the reference's untouched sum[i]
    double sumsq1 = dSub(q0, qm);
    float sum2 = d2f(dSub(sp, s0));
    float sumsq2 = d2f(dSub(qp, q0));
    double wd = (double)wf;
    float mean1 = d2f(dDiv(sum1, wd));
    float mean2 = fDiv(sum2, wf);
    float m1sq = fMul(mean1, mean1);
    float q2 = fDiv(sumsq2, wf);
    float m2sq = fMul(mean2, mean2);
    float cv = d2f(dSub(dAdd(dSub(dDiv(sumsq1, wd), (double)m1sq), (double)q2), (double)m2sq));
    cv = fmaxf(cv, FLT_MIN);
    float dm = fSub(mean2, mean1);
    return d2f(dDiv(fabs((double)dm), __dsqrt_rn((double)fDiv(cv, wf))));
}

template <bool kI16>
__global__ void __launch_bounds__(SEG_THREADS) seg_kernel(DnbBatchView v, DnbDetector det) {
    __shared__ double ring_s[SEG_RING][SEG_THREADS];
    __shared__ double ring_q[SEG_RING][SEG_THREADS];
    const int tid = threadIdx.x;
    const uint32_t slot = blockIdx.x * SEG_THREADS + tid;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    const int w1 = (int)det.w1, w2 = (int)det.w2;
    const float w1f = (float)det.w1, w2f = (float)det.w2;
    const bool t1_on = (N >= 2u * det.w1) && det.w1 >= 2, t2_on = (N >= 2u * det.w2) && det.w2 >= 2;
    float dac_off = 0.f, dac_scl = 1.f;
    if (kI16) { dac_off = v.dac_offset[r]; dac_scl = v.dac_scale[r]; }

    EventSink sink;
    sink.prev_peak = 0; sink.prev_sum = 0.0; sink.prev_sq = 0.0; sink.et_count = 0;
    sink.E = 0; sink.pend_mean = 0.0f; sink.pend_start = 0;
    sink.cap = (uint32_t)(v.ev_off[r + 1] - v.ev_off[r]);
    sink.ev_start = v.ev_start + v.ev_off[r] + r;
    sink.ev_mean = v.ev_mean + v.ev_off[r];
    sink.et_start = v.et_start ? v.et_start + v.ev_off[r] + r : nullptr;
    sink.et_length = v.et_start ? v.et_length + v.ev_off[r] + r : nullptr;
    sink.et_mean = v.et_start ? v.et_mean + v.ev_off[r] + r : nullptr;
    sink.et_stdv = v.et_start ? v.et_stdv + v.ev_off[r] + r : nullptr;

    Detector ds = {-1, FLT_MAX, false, 0.0, 0.0}, dl = {-1, FLT_MAX, false, 0.0, 0.0};
    uint32_t l_masked_to = 0;
    double sum = 0.0, sumsq = 0.0;
    ring_s[0][tid] = 0.0;
    ring_q[0][tid] = 0.0;

    // one position of both detectors (event_detection.c:136-194); t1/t2 are the t-statistics at i
    auto fsm = [&](int i, float t1, float t2) {
        if (i > 0) {   // short detector: masked_to stays 0, so only i == 0 is skipped (:140)
            float cur = t1;
            if (ds.pos == -1) {
                if (cur < ds.val) {
                    ds.val = cur;
                } else if (fSub(cur, ds.val) > det.peak_height) {
                    ds.val = cur; ds.pos = i;
                    ds.psum = ring_s[i & (SEG_RING - 1)][tid]; ds.psq = ring_q[i & (SEG_RING - 1)][tid];
                }
            } else {
                if (cur > ds.val) {
                    ds.val = cur; ds.pos = i;
                    ds.psum = ring_s[i & (SEG_RING - 1)][tid]; ds.psq = ring_q[i & (SEG_RING - 1)][tid];
                }
                if (ds.val > det.thr1) {   // :165-176 the short detector dominates the long one
                    l_masked_to = (uint32_t)ds.pos + det.w1;
                    dl.pos = -1; dl.val = FLT_MAX; dl.valid = false;
                }
                if (fSub(ds.val, cur) > det.peak_height && ds.val > det.thr1) ds.valid = true;
                if (ds.valid && (uint32_t)(i - ds.pos) > det.w1 / 2) {
                    finish_event(sink, (uint32_t)ds.pos, ds.psum, ds.psq);
                    ds.pos = -1; ds.val = cur; ds.valid = false;
                }
            }
        }
        if (!(l_masked_to >= (uint32_t)i)) {
            float cur = t2;
            if (dl.pos == -1) {
                if (cur < dl.val) {
                    dl.val = cur;
                } else if (fSub(cur, dl.val) > det.peak_height) {
                    dl.val = cur; dl.pos = i;
                    dl.psum = ring_s[i & (SEG_RING - 1)][tid]; dl.psq = ring_q[i & (SEG_RING - 1)][tid];
                }
            } else {
                if (cur > dl.val) {
                    dl.val = cur; dl.pos = i;
                    dl.psum = ring_s[i & (SEG_RING - 1)][tid]; dl.psq = ring_q[i & (SEG_RING - 1)][tid];
                }
                if (fSub(dl.val, cur) > det.peak_height && dl.val > det.thr2) dl.valid = true;
                if (dl.valid && (uint32_t)(i - dl.pos) > det.w2 / 2) {
                    finish_event(sink, (uint32_t)dl.pos, dl.psum, dl.psq);
                    dl.pos = -1; dl.val = cur; dl.valid = false;
                }
            }
        }
    };

    auto position = [&](int i) {
        float t1 = 0.0f, t2 = 0.0f;
        const double s0 = ring_s[i & (SEG_RING - 1)][tid], q0 = ring_q[i & (SEG_RING - 1)][tid];
        if (t1_on && i >= w1 && (uint32_t)i <= N - det.w1)
            t1 = tstat_at(ring_s[(i - w1) & (SEG_RING - 1)][tid], ring_q[(i - w1) & (SEG_RING - 1)][tid], s0, q0,
                          ring_s[(i + w1) & (SEG_RING - 1)][tid], ring_q[(i + w1) & (SEG_RING - 1)][tid], w1f);
        if (t2_on && i >= w2 && (uint32_t)i <= N - det.w2)
            t2 = tstat_at(ring_s[(i - w2) & (SEG_RING - 1)][tid], ring_q[(i - w2) & (SEG_RING - 1)][tid], s0, q0,
                          ring_s[(i + w2) & (SEG_RING - 1)][tid], ring_q[(i + w2) & (SEG_RING - 1)][tid], w2f);
        fsm(i, t1, t2);
    };

    auto consume = [&](int j, float xf) {
        double x = (double)xf;
        sum = dAdd(sum, x);                 // event_detection.c:45
        sumsq = dAdd(sumsq, dMul(x, x));    // :46 (x*x is exact for a float-valued x)
        ring_s[(j + 1) & (SEG_RING - 1)][tid] = sum;
        ring_q[(j + 1) & (SEG_RING - 1)][tid] = sumsq;
        int i = j + 1 - w2;
        if (i >= 0) position(i);
    };

    // stream the signal, 4 samples per load, next load in flight while the current four are consumed
    const uint32_t n4 = N & ~3u;
    if (kI16) {
        const short4 *p = reinterpret_cast<const short4 *>(v.raw_i16 + base);
        short4 cur = n4 ? __ldg(p) : make_short4(0, 0, 0, 0);
        for (uint32_t j = 0; j < n4; j += 4) {
            short4 nxt = (j + 4 < n4) ? __ldg(p + (j >> 2) + 1) : make_short4(0, 0, 0, 0);
            consume((int)j + 0, fMul(fAdd((float)cur.x, dac_off), dac_scl));   // src/pod5.cpp:60
            consume((int)j + 1, fMul(fAdd((float)cur.y, dac_off), dac_scl));
            consume((int)j + 2, fMul(fAdd((float)cur.z, dac_off), dac_scl));
            consume((int)j + 3, fMul(fAdd((float)cur.w, dac_off), dac_scl));
            cur = nxt;
        }
        for (uint32_t j = n4; j < N; j++) consume((int)j, fMul(fAdd((float)v.raw_i16[base + j], dac_off), dac_scl));
    } else {
        const float4 *p = reinterpret_cast<const float4 *>(v.raw_f32 + base);
        float4 cur = n4 ? __ldg(p) : make_float4(0, 0, 0, 0);
        for (uint32_t j = 0; j < n4; j += 4) {
            float4 nxt = (j + 4 < n4) ? __ldg(p + (j >> 2) + 1) : make_float4(0, 0, 0, 0);
            consume((int)j + 0, cur.x);
            consume((int)j + 1, cur.y);
            consume((int)j + 2, cur.z);
            consume((int)j + 3, cur.w);
            cur = nxt;
        }
        for (uint32_t j = n4; j < N; j++) consume((int)j, v.raw_f32[base + j]);
    }
    // tail positions whose look-ahead runs past the end (t2 == 0 there)
    {
        int i0 = (int)N - w2 + 1;
        if (i0 < 0) i0 = 0;
        for (int i = i0; i < (int)N; i++) position(i);
    }

    int status = DNB_READ_OK;
    if (sink.et_count == 0) {
        // no peak: the reference indexes peaks[n-2] with n == 1 (event_detection.c:263) -- undefined
        v.et_n[r] = 0;
        v.n_events[r] = 0;
        status = DNB_READ_UNDEFINED;
    } else {
        finish_event(sink, N, sum, sumsq);      // last event ends at nsample (:263)
        if (sink.E <= sink.cap) sink.ev_start[sink.E] = sink.pend_start;
        v.et_n[r] = sink.et_count;
        v.n_events[r] = sink.E;
        if (sink.E > sink.cap || (v.et_start && sink.et_count > sink.cap + 1)) status = DNB_READ_OVERFLOW;
        else if (sink.E == 0) status = DNB_READ_UNDEFINED;
    }
    if (!v.et_start) {
        // query no longer than k / reference shorter than k: n_kmers and eventsPerBase are degenerate in the reference
        const uint64_t ql = v.q_off[r + 1] - v.q_off[r], rl = v.r_off[r + 1] - v.r_off[r];
        if (ql <= DNB_K || rl < DNB_K) status = DNB_READ_UNDEFINED;
    }
    v.status[r] = status;
}

}  // namespace

void dnb_launch_segmentation(const DnbBatchView &v, DnbDetector det, cudaStream_t s) {
    if (v.n_reads == 0) return;
    dim3 grid((v.n_reads + SEG_THREADS - 1) / SEG_THREADS);
    if (v.raw_i16)
        seg_kernel<true><<<grid, SEG_THREADS, 0, s>>>(v, det);
    else
        seg_kernel<false><<<grid, SEG_THREADS, 0, s>>>(v, det);
}
