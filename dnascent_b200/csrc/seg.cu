// seg.cu -- scrappie t-statistic event segmentation + the r.events filter.
//
// Replaces (reference, paths relative to /root/reference):
//   compute_sum_sumsq            src/scrappie/event_detection.c:35-48
//   compute_tstat (w=3, w=6)     src/scrappie/event_detection.c:60-115
//   short_long_peak_detector     src/scrappie/event_detection.c:122-198
//   create_event(s)              src/scrappie/event_detection.c:213-266
//   event table -> r.events      src/event_handling.cpp:549-575   (quirks Q1-Q3)
//
// Exactness constraints.  (1) The double prefix sum of squares rounds at EVERY step (x*x needs up to 48 bits, the
// running sum passes 2^19 within ~100 samples), so its value depends on the serial order; the survey measured 10 %
// of reads changing boundaries under a different summation.  (2) The two peak detectors are a sequential state
// machine over the samples.  Both are made parallel here without giving up a single bit:
//
//   K1 checkpoint   one lane per read runs ONLY the two serial double chains (3 flops per sample) and stores
//                   (sum, sumsq) every 64 samples.
//   K2 tiles        one lane per 512-sample tile restarts the chains from the exact checkpoint, streams both
//                   t-statistics from a 16-deep shared-memory ring and runs the detector pair SPECULATIVELY from a
//                   fresh state 64 samples before its tile; peaks emitted inside the tile go to a per-tile list.
//                   The detector state it assumed at the tile start and the state it ends with are recorded.
//   K3 stitch       per read: a tile's assumed start state must equal its predecessor's end state (tile 0 starts
//                   from the true initial state), which by induction makes every emitted peak the serial one;
//                   otherwise the read is flagged.  Also prefix-counts the peaks.
//   K4 events       one thread per peak builds the scrappie event ending there and the r.events arrays.
//   K5 serial       flagged reads (state mismatch, list overflow, a non-positive event mean -- all rare) are redone by
//                   the one-lane-per-read serial kernel, which is a literal transcription and also serves
//                   dnb_detect_events (full event table with stdv).
//
// The signal is read twice (K1, K2 + halo) and only the event table is written: algorithmic traffic is
// 4 B/sample + 8 B/event (2 B/sample for int16 DAC input).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define SEG_THREADS 64
#define SEG_RING 16
#ifndef SEG_TILE_MIN_BLOCKS
#define SEG_TILE_MIN_BLOCKS 12     // resident CTAs per SM the tile kernel's register allocation must allow: 80 registers
                                   // (24 warps per SM, 24 B of spills) against 92 unconstrained (20 warps): 125 -> 116 ms per 30k-read step
#endif

namespace {

struct Detector {
    int pos;          // peak_pos, -1 = DEF_PEAK_POS
    float val;        // peak_value
    bool valid;       // valid_peak
    double psum, psq; // sums[peak_pos], sumsqs[peak_pos] (carried so that create_event needs no array)
};

// ---- sinks: what happens to an emitted peak -------------------------------------------------------------------
// serial kernel: events are built on the fly (event_detection.c:213-266 + event_handling.cpp:549-575)
struct EventSink {
    uint32_t prev_peak;
    double prev_sum, prev_sq;
    uint32_t et_count;
    uint32_t E;
    float pend_mean;
    uint32_t pend_start;
    uint32_t cap;
    uint32_t *ev_start;
    float *ev_mean;
    uint64_t *et_start;
    float *et_length, *et_mean, *et_stdv;

    __device__ __forceinline__ void finish_event(uint32_t end, double esum, double esq) {
        unsigned long long span = (unsigned long long)((long long)end - (long long)prev_peak);  // size_t arithmetic
        float length = (float)span;
        float mean = fDiv(d2f(dSub(esum, prev_sum)), length);
        uint32_t idx = et_count++;
        if (et_start && idx <= cap) {
            float dsq = d2f(dSub(esq, prev_sq));
            float var = fSub(fDiv(dsq, length), fMul(mean, mean));
            et_start[idx] = prev_peak;
            et_length[idx] = length;
            et_mean[idx] = mean;
            et_stdv[idx] = __fsqrt_rn(fmaxf(var, 0.0f));
        }
        if ((double)mean > 0.0 && idx > 0) {
            if (E < cap) {
                ev_mean[E] = pend_mean;
                ev_start[E] = pend_start;
            }
            E++;
            pend_mean = mean;
            pend_start = prev_peak;
        }
        prev_peak = end;
        prev_sum = esum;
        prev_sq = esq;
    }
    __device__ __forceinline__ void on_peak(int /*i_emit*/, int pos, double psum, double psq) {
        finish_event((uint32_t)pos, psum, psq);
    }
};

// tile kernel: peaks emitted at a time inside the tile are listed
struct PeakSink {
    int t0;
    uint32_t count, cap;
    uint32_t *pos_out;
    double *sum_out;
    __device__ __forceinline__ void on_peak(int i_emit, int pos, double psum, double /*psq*/) {
        if (i_emit < t0) return;   // emitted while still warming up in the halo: belongs to the previous tile
        if (count < cap) {
            pos_out[count] = (uint32_t)pos;
            sum_out[count] = psum;
        }
        count++;
    }
};

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

// ---- the same t-statistic without IEEE division / square-root subroutines (tile kernel only) ----------------------
// Bit-identical to tstat_at for every read that passes the sample precondition checked by the checkpoint kernel
// (every sample 0 or 2^-20 <= |x| < 2^30; other reads go to the literal serial kernel):
//  * x / w with w a small integer constant: q0 = x*r, rem = fma(-q0, w, x), q = fma(rem, r, q0), r = RN(1/w)
//    (Markstein's correction step).  It returns the correctly rounded quotient when |1 - w*r| <= 2^-54 (double;
//    checked on the host for the configured windows, DnbDetector::fast_div) resp. for every finite float with
//    |x| >= 2^-124 and w = 2..7 (checked exhaustively, tests/test_host_logic.py::test_constant_divisor_sequences).
//  * |dm| / sqrt(v): rsqrt.approx (2^-22) + one cubic correction gives |dm| * v^-1/2 to < 8 ulp of the reference's
//    RN(|dm| / RN(sqrt v)); if that double lies within 64 ulp of a float rounding boundary, or the seed was worse than
//    2^-20, or cv is near the subnormal range, the lane recomputes the tail with the IEEE operations.
struct FastDiv {
    double wd, rd;
    float wf, rf;
};
__device__ __forceinline__ double div_const(double x, const FastDiv &f) {
    const double q0 = dMul(x, f.rd);
    const double rem = __fma_rn(-q0, f.wd, x);      // exact remainder
    return __fma_rn(rem, f.rd, q0);
}
__device__ __forceinline__ float div_const(float x, const FastDiv &f) {
    const float q0 = fMul(x, f.rf);
    const float rem = __fmaf_rn(-q0, f.wf, x);
    return __fmaf_rn(rem, f.rf, q0);
}
__device__ __forceinline__ double rsqrt_seed(double v) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    return y;
}
__device__ __forceinline__ float tstat_fast(double sm, double qm, double s0, double q0, double sp, double qp, const FastDiv &f) {
    const double sum1 = dSub(s0, sm);
    const double sumsq1 = dSub(q0, qm);
    const float sum2 = d2f(dSub(sp, s0));
    const float sumsq2 = d2f(dSub(qp, q0));
    const float mean1 = d2f(div_const(sum1, f));
    const float mean2 = div_const(sum2, f);
    const float m1sq = fMul(mean1, mean1);
    const float q2 = div_const(sumsq2, f);
    const float m2sq = fMul(mean2, mean2);
    float cv = d2f(dSub(dAdd(dSub(div_const(sumsq1, f), (double)m1sq), (double)q2), (double)m2sq));
    cv = fmaxf(cv, FLT_MIN);
    const double adm = fabs((double)fSub(mean2, mean1));
    const double v = (double)div_const(cv, f);
    const double y = rsqrt_seed(v);
    const double e = __fma_rn(-dMul(v, y), y, 1.0);                    // 1 - v*y^2
    const double y1 = __fma_rn(dMul(y, e), __fma_rn(e, 0.375, 0.5), y);   // y * (1 + e/2 + 3e^2/8)
    const double t = dMul(adm, y1);
    bool slow = !(cv >= 7.888609052210118e-31f);                           // 2^-100 (also catches NaN)
    slow |= ((uint32_t)__double2hiint(e) & 0x7ff00000u) > 0x3eb00000u;    // |e| >= 2^-20: seed not as accurate as assumed
    slow |= (((uint32_t)__double2loint(t) & 0x1fffffffu) - (0x10000000u - 64u)) <= 128u;
    if (slow) return d2f(dDiv(adm, __dsqrt_rn((double)fDiv(cv, f.wf))));
    return d2f(t);
}

__device__ __forceinline__ bool same_boundary(const SegBoundary &a, const SegBoundary &b) {
    return a.s_pos == b.s_pos && a.l_pos == b.l_pos && __float_as_uint(a.s_val) == __float_as_uint(b.s_val) &&
           __float_as_uint(a.l_val) == __float_as_uint(b.l_val) && a.masked == b.masked && a.valid == b.valid;
}

// ---- the streaming engine: serial chains + t-statistics + detector pair over positions [first_pos, end_pos) ------
template <class Sink, bool kFast = false>
struct SegEngine {
    uint32_t N;
    DnbDetector det;
    int w1, w2;
    float w1f, w2f;
    FastDiv f1, f2;
    bool t1_on, t2_on;
    double *rs, *rq;            // this thread's column of the shared-memory rings (stride SEG_THREADS)
    Detector ds, dl;
    uint32_t l_masked_to;
    double sum, sumsq;
    int next_pos, end_pos;      // positions still to run through the detectors: [next_pos, end_pos)
    int capture_at;             // position before which the detector state is snapshotted (-1: never)
    SegBoundary snap;
    Sink sink;

    __device__ __forceinline__ double &RS(int i) { return rs[(i & (SEG_RING - 1)) * SEG_THREADS]; }
    __device__ __forceinline__ double &RQ(int i) { return rq[(i & (SEG_RING - 1)) * SEG_THREADS]; }

    __device__ void init(uint32_t n, const DnbDetector &d, double *ring_s, double *ring_q, int chain_start, double s0, double q0) {
        N = n; det = d; w1 = (int)d.w1; w2 = (int)d.w2; w1f = (float)d.w1; w2f = (float)d.w2;
        f1 = {(double)d.w1, dDiv(1.0, (double)d.w1), w1f, fDiv(1.0f, w1f)};
        f2 = {(double)d.w2, dDiv(1.0, (double)d.w2), w2f, fDiv(1.0f, w2f)};
        t1_on = (n >= 2u * d.w1) && d.w1 >= 2; t2_on = (n >= 2u * d.w2) && d.w2 >= 2;
        rs = ring_s; rq = ring_q;
        ds = {-1, FLT_MAX, false, 0.0, 0.0}; dl = {-1, FLT_MAX, false, 0.0, 0.0};
        l_masked_to = 0;
        sum = s0; sumsq = q0;
        RS(chain_start) = s0; RQ(chain_start) = q0;
        capture_at = -1;
    }

    __device__ __forceinline__ SegBoundary boundary(int p) const {
        SegBoundary b;
        b.s_pos = ds.pos; b.l_pos = dl.pos; b.s_val = ds.val; b.l_val = dl.val;
        b.masked = (l_masked_to >= (uint32_t)p) ? l_masked_to : 0u;
        b.valid = (ds.valid ? 1u : 0u) | (dl.valid ? 2u : 0u);
        return b;
    }

    // one position of both detectors (event_detection.c:136-194); t1/t2 are the t-statistics at i
    __device__ __forceinline__ void fsm(int i, float t1, float t2) {
        if (i > 0) {   // short detector: masked_to stays 0, so only i == 0 is skipped (:140)
            const float cur = t1;
            if (ds.pos == -1) {
                if (cur < ds.val) {
                    ds.val = cur;
                } else if (fSub(cur, ds.val) > det.peak_height) {
                    ds.val = cur; ds.pos = i; ds.psum = RS(i); ds.psq = RQ(i);
                }
            } else {
                if (cur > ds.val) { ds.val = cur; ds.pos = i; ds.psum = RS(i); ds.psq = RQ(i); }
                if (ds.val > det.thr1) {   // :165-176 the short detector dominates the long one
                    l_masked_to = (uint32_t)ds.pos + det.w1;
                    dl.pos = -1; dl.val = FLT_MAX; dl.valid = false;
                }
                if (fSub(ds.val, cur) > det.peak_height && ds.val > det.thr1) ds.valid = true;
                if (ds.valid && (uint32_t)(i - ds.pos) > det.w1 / 2) {
                    sink.on_peak(i, ds.pos, ds.psum, ds.psq);
                    ds.pos = -1; ds.val = cur; ds.valid = false;
                }
            }
        }
        if (!(l_masked_to >= (uint32_t)i)) {
            const float cur = t2;
            if (dl.pos == -1) {
                if (cur < dl.val) {
                    dl.val = cur;
                } else if (fSub(cur, dl.val) > det.peak_height) {
                    dl.val = cur; dl.pos = i; dl.psum = RS(i); dl.psq = RQ(i);
                }
            } else {
                if (cur > dl.val) { dl.val = cur; dl.pos = i; dl.psum = RS(i); dl.psq = RQ(i); }
                if (fSub(dl.val, cur) > det.peak_height && dl.val > det.thr2) dl.valid = true;
                if (dl.valid && (uint32_t)(i - dl.pos) > det.w2 / 2) {
                    sink.on_peak(i, dl.pos, dl.psum, dl.psq);
                    dl.pos = -1; dl.val = cur; dl.valid = false;
                }
            }
        }
    }

    __device__ __forceinline__ void position(int i) {
        if (i == capture_at) snap = boundary(i);
        float t1 = 0.0f, t2 = 0.0f;
        const double s0 = RS(i), q0 = RQ(i);
        if (t1_on && i >= w1 && (uint32_t)i <= N - det.w1)
            t1 = kFast ? tstat_fast(RS(i - w1), RQ(i - w1), s0, q0, RS(i + w1), RQ(i + w1), f1)
                       : tstat_at(RS(i - w1), RQ(i - w1), s0, q0, RS(i + w1), RQ(i + w1), w1f);
        if (t2_on && i >= w2 && (uint32_t)i <= N - det.w2)
            t2 = kFast ? tstat_fast(RS(i - w2), RQ(i - w2), s0, q0, RS(i + w2), RQ(i + w2), f2)
                       : tstat_at(RS(i - w2), RQ(i - w2), s0, q0, RS(i + w2), RQ(i + w2), w2f);
        fsm(i, t1, t2);
    }

    __device__ __forceinline__ void consume(int j, float xf) {
        const double x = (double)xf;
        sum = dAdd(sum, x);                 // event_detection.c:45
        sumsq = dAdd(sumsq, dMul(x, x));    // :46 (x*x is exact for a float-valued x)
        RS(j + 1) = sum;
        RQ(j + 1) = sumsq;
        const int i = j + 1 - w2;           // newest position whose look-ahead is complete
        if (i >= next_pos && next_pos < end_pos) { position(next_pos); next_pos++; }
    }

    // positions whose look-ahead runs past the end of the read (their t2 is 0 by definition)
    __device__ __forceinline__ void drain() {
        while (next_pos < end_pos) { position(next_pos); next_pos++; }
    }
};

template <bool kI16>
struct SampleReader {
    const float *f32;
    const int16_t *i16;
    float dac_off, dac_scl;
    __device__ __forceinline__ float one(uint64_t idx) const {
        if (kI16) return fMul(fAdd((float)i16[idx], dac_off), dac_scl);   // src/pod5.cpp:60
        return f32[idx];
    }
    // four consecutive samples starting at a multiple of 4
    __device__ __forceinline__ void four(uint64_t idx, float o[4]) const {
        if (kI16) {
            const short4 s = __ldg(reinterpret_cast<const short4 *>(i16 + idx));
            o[0] = fMul(fAdd((float)s.x, dac_off), dac_scl); o[1] = fMul(fAdd((float)s.y, dac_off), dac_scl);
            o[2] = fMul(fAdd((float)s.z, dac_off), dac_scl); o[3] = fMul(fAdd((float)s.w, dac_off), dac_scl);
        } else {
            const float4 s = __ldg(reinterpret_cast<const float4 *>(f32 + idx));
            o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
        }
    }
};

template <bool kI16, class Engine>
__device__ __forceinline__ void stream(Engine &en, const SampleReader<kI16> &rd, uint64_t base, int j0, int j1) {
    // j0 is a multiple of 4 and base + j0 is 16-byte aligned (read starts are 32-element aligned)
    const int j4 = j0 + ((j1 - j0) & ~3);
    float cur[4], nxt[4] = {0.f, 0.f, 0.f, 0.f};
    if (j0 < j4) rd.four(base + j0, cur);
    for (int j = j0; j < j4; j += 4) {
        if (j + 4 < j4) rd.four(base + j + 4, nxt);
        en.consume(j + 0, cur[0]);
        en.consume(j + 1, cur[1]);
        en.consume(j + 2, cur[2]);
        en.consume(j + 3, cur[3]);
        cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2]; cur[3] = nxt[3];
    }
    for (int j = j4; j < j1; j++) en.consume(j, rd.one(base + j));
}

// ---- K5 / dnb_detect_events: literal serial transcription, one lane per read ----------------------------------------
template <bool kI16>
__global__ void __launch_bounds__(SEG_THREADS) seg_serial_kernel(DnbBatchView v, DnbDetector det, const uint32_t *only_flagged) {
    __shared__ double ring_s[SEG_RING][SEG_THREADS];
    __shared__ double ring_q[SEG_RING][SEG_THREADS];
    const int tid = threadIdx.x;
    const uint32_t slot = blockIdx.x * SEG_THREADS + tid;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    if (only_flagged && !only_flagged[r]) return;
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    SampleReader<kI16> rd{v.raw_f32, v.raw_i16, 0.f, 1.f};
    if (kI16) { rd.dac_off = v.dac_offset[r]; rd.dac_scl = v.dac_scale[r]; }

    SegEngine<EventSink> en;
    en.init(N, det, &ring_s[0][tid], &ring_q[0][tid], 0, 0.0, 0.0);
    en.next_pos = 0; en.end_pos = (int)N;
    EventSink &sink = en.sink;
    sink.prev_peak = 0; sink.prev_sum = 0.0; sink.prev_sq = 0.0; sink.et_count = 0;
    sink.E = 0; sink.pend_mean = 0.0f; sink.pend_start = 0;
    sink.cap = (uint32_t)(v.ev_off[r + 1] - v.ev_off[r]);
    sink.ev_start = v.ev_start + v.ev_off[r] + r;
    sink.ev_mean = v.ev_mean + v.ev_off[r];
    sink.et_start = v.et_start ? v.et_start + v.ev_off[r] + r : nullptr;
    sink.et_length = v.et_start ? v.et_length + v.ev_off[r] + r : nullptr;
    sink.et_mean = v.et_start ? v.et_mean + v.ev_off[r] + r : nullptr;
    sink.et_stdv = v.et_start ? v.et_stdv + v.ev_off[r] + r : nullptr;

    stream<kI16>(en, rd, base, 0, (int)N);
    en.drain();

    int status = DNB_READ_OK;
    if (sink.et_count == 0) {
        // no peak: the reference indexes peaks[n-2] with n == 1 (event_detection.c:263) -- undefined
        v.et_n[r] = 0;
        v.n_events[r] = 0;
        status = DNB_READ_UNDEFINED;
    } else {
        sink.finish_event(N, en.sum, en.sumsq);      // last event ends at nsample (:263)
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

// ---- K1: exact (sum, sumsq) checkpoints every DNB_SEG_CK samples, one lane per read ------------------------------------
// The two chains cost 3 flops per sample but are serial, so this kernel is a latency machine: a lane may not wait
// for memory.  Samples come in 16-byte vectors, one group (32 samples) ahead in registers, and the lines of the
// groups after that are pulled into L2 with prefetch instructions ~2 KB ahead of the chain.
#define CK_GROUP 32   // samples per register group

template <bool kI16>
struct CkGroup {
    uint4 q[kI16 ? CK_GROUP / 8 : CK_GROUP / 4];
    __device__ __forceinline__ void load(const SampleReader<kI16> &rd, uint64_t idx) {
        const uint4 *p = kI16 ? reinterpret_cast<const uint4 *>(rd.i16 + idx) : reinterpret_cast<const uint4 *>(rd.f32 + idx);
#pragma unroll
        for (int i = 0; i < (kI16 ? CK_GROUP / 8 : CK_GROUP / 4); i++) q[i] = __ldg(p + i);
    }
    __device__ __forceinline__ float get(const SampleReader<kI16> &rd, int i) const {
        if (kI16) {
            const uint32_t w = (&q[i >> 3].x)[(i >> 1) & 3];
            const short sv = (short)((i & 1) ? (w >> 16) : (w & 0xffffu));
            return fMul(fAdd((float)sv, rd.dac_off), rd.dac_scl);                       // src/pod5.cpp:60
        }
        return __uint_as_float((&q[i >> 2].x)[i & 3]);
    }
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <bool kI16>
__global__ void __launch_bounds__(128) seg_checkpoint_kernel(DnbBatchView v, DnbSegTiles t) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    SampleReader<kI16> rd{v.raw_f32, v.raw_i16, 0.f, 1.f};
    if (kI16) { rd.dac_off = v.dac_offset[r]; rd.dac_scl = v.dac_scale[r]; }
    double *cs = t.ck_sum + t.ck_off[r], *cq = t.ck_sq + t.ck_off[r];
    const char *bytes = kI16 ? reinterpret_cast<const char *>(rd.i16 + base) : reinterpret_cast<const char *>(rd.f32 + base);
    const uint32_t group_bytes = CK_GROUP * (kI16 ? 2 : 4);
    const uint64_t total_bytes = (uint64_t)N * (kI16 ? 2 : 4);
    const uint32_t pf_ahead = 2048;
    double sum = 0.0, sumsq = 0.0;
    uint32_t out_of_range = 0;     // precondition of tstat_fast: every sample is 0 or 2^-20 <= |x| < 2^30
    const uint32_t ng = N / CK_GROUP;
    for (uint32_t o = 0; o < pf_ahead && o < total_bytes; o += 128) prefetch_l2(bytes + o);
    CkGroup<kI16> cur, nxt;
    if (ng) cur.load(rd, base);
    for (uint32_t g = 0; g < ng; g++) {
        if (g + 1 < ng) nxt.load(rd, base + (uint64_t)(g + 1) * CK_GROUP);
        {
            const uint64_t o = (uint64_t)g * group_bytes + pf_ahead;
            if ((o & 127) == 0 && o < total_bytes) prefetch_l2(bytes + o);
        }
        const uint32_t j = g * CK_GROUP;
        if ((j & (DNB_SEG_CK - 1)) == 0) { cs[j / DNB_SEG_CK] = sum; cq[j / DNB_SEG_CK] = sumsq; }
#pragma unroll
        for (int u = 0; u < CK_GROUP; u++) {
            const float xf = cur.get(rd, u);
            const double x = (double)xf;
            const uint32_t ax = __float_as_uint(xf) & 0x7fffffffu;
            out_of_range |= (ax - 0x35800000u >= 0x19000000u) && ax != 0u;
            sum = dAdd(sum, x);                  // event_detection.c:45
            sumsq = dAdd(sumsq, dMul(x, x));     // :46 (x*x is exact for a float-valued x)
        }
        cur = nxt;
    }
    for (uint32_t j = ng * CK_GROUP; j < N; j++) {
        if ((j & (DNB_SEG_CK - 1)) == 0) { cs[j / DNB_SEG_CK] = sum; cq[j / DNB_SEG_CK] = sumsq; }
        const float xf = rd.one(base + j);
        const double x = (double)xf;
        const uint32_t ax = __float_as_uint(xf) & 0x7fffffffu;
        out_of_range |= (ax - 0x35800000u >= 0x19000000u) && ax != 0u;
        sum = dAdd(sum, x);
        sumsq = dAdd(sumsq, dMul(x, x));
    }
    t.tot_sum[r] = sum;
    t.redo[r] = out_of_range;      // the stitch kernel ORs its own verdict into this
}

// ---- K2: one lane per tile ------------------------------------------------------------------------------------------
template <bool kI16, bool kFast>
__global__ void __launch_bounds__(SEG_THREADS, SEG_TILE_MIN_BLOCKS) seg_tile_kernel(DnbBatchView v, DnbDetector det, DnbSegTiles t) {
    __shared__ double ring_s[SEG_RING][SEG_THREADS];
    __shared__ double ring_q[SEG_RING][SEG_THREADS];
    const int tid = threadIdx.x;
    const uint32_t g = blockIdx.x * SEG_THREADS + tid;
    if (g >= t.n_tiles) return;
    const uint32_t r = t.tile_read[g];
    const uint32_t tile = g - t.tile_off[r];
    const uint32_t N = v.n_samples[r];
    const uint64_t base = v.raw_off[r];
    SampleReader<kI16> rd{v.raw_f32, v.raw_i16, 0.f, 1.f};
    if (kI16) { rd.dac_off = v.dac_offset[r]; rd.dac_scl = v.dac_scale[r]; }
    const int w2 = (int)det.w2;
    const int t0 = (int)(tile * DNB_SEG_TILE);
    const int t1 = (int)min((uint32_t)t0 + DNB_SEG_TILE, N);
    const int s0 = tile == 0 ? 0 : t0 - DNB_SEG_HALO;                       // detectors start here, fresh
    const int c0 = tile == 0 ? 0 : ((s0 - w2) / DNB_SEG_CK) * DNB_SEG_CK;   // chains restart at this checkpoint
    const uint64_t ck = t.ck_off[r] + (uint32_t)c0 / DNB_SEG_CK;

    SegEngine<PeakSink, kFast> en;
    en.init(N, det, &ring_s[0][tid], &ring_q[0][tid], c0, t.ck_sum[ck], t.ck_sq[ck]);
    en.next_pos = s0; en.end_pos = t1;
    en.capture_at = tile == 0 ? -1 : t0;
    en.sink.t0 = t0; en.sink.count = 0; en.sink.cap = DNB_SEG_PEAK_CAP;
    en.sink.pos_out = t.pk_pos + (size_t)g * DNB_SEG_PEAK_CAP;
    en.sink.sum_out = t.pk_sum + (size_t)g * DNB_SEG_PEAK_CAP;
    en.snap = en.boundary(0);

    const int j1 = (int)min((uint32_t)(t1 + w2 - 1), N);   // last sample needed so that position t1-1 has its look-ahead
    stream<kI16>(en, rd, base, c0, j1);
    if ((uint32_t)j1 == N) en.drain();                       // tile touches the end of the read

    t.pk_count[g] = en.sink.count;
    t.b_start[g] = en.snap;
    t.b_end[g] = en.boundary(t1);
}

// ---- K3: stitch the tiles of a read, one warp per read ---------------------------------------------------------------
__global__ void __launch_bounds__(128) seg_stitch_kernel(DnbBatchView v, DnbSegTiles t) {
    const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= v.n_reads) return;
    const uint32_t r = v.order[slot];
    const uint32_t g0 = t.tile_off[r], g1 = t.tile_off[r + 1];
    bool bad = false;
    uint32_t prefix = 0, last_pos = 0;
    double last_sum = 0.0;
    for (uint32_t gb = g0; gb < g1 && !bad; gb += 32) {
        const uint32_t g = gb + lane;
        const bool ok = g < g1;
        uint32_t c = 0;
        bool mism = false;
        if (ok) {
            c = t.pk_count[g];
            if (g > g0 && !same_boundary(t.b_start[g], t.b_end[g - 1])) mism = true;
            if (c > DNB_SEG_PEAK_CAP) mism = true;
        }
        if (__any_sync(0xffffffffu, mism)) { bad = true; break; }
        // exclusive prefix of the peak counts; last peak (pos, sum) before each tile = that of the nearest earlier
        // tile with a peak
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const unsigned has = __ballot_sync(0xffffffffu, ok && c > 0);
        uint32_t my_pos = 0;
        double my_sum = 0.0;
        if (ok && c > 0) {
            my_pos = t.pk_pos[(size_t)g * DNB_SEG_PEAK_CAP + c - 1];
            my_sum = t.pk_sum[(size_t)g * DNB_SEG_PEAK_CAP + c - 1];
        }
        const unsigned before = has & ((1u << lane) - 1u);
        const int src = before ? 31 - __clz(before) : 0;
        const uint32_t p_pos = __shfl_sync(0xffffffffu, my_pos, src);
        const double p_sum = __shfl_sync(0xffffffffu, my_sum, src);
        if (ok) {
            t.tile_prefix[g] = prefix + incl - c;
            t.tile_prev_pos[g] = before ? p_pos : last_pos;
            t.tile_prev_sum[g] = before ? p_sum : last_sum;
        }
        if (has) {
            const int top = 31 - __clz(has);
            last_pos = __shfl_sync(0xffffffffu, my_pos, top);
            last_sum = __shfl_sync(0xffffffffu, my_sum, top);
        }
        prefix += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane != 0) return;
    const uint32_t cap = (uint32_t)(v.ev_off[r + 1] - v.ev_off[r]);
    int status = DNB_READ_OK;
    if (t.redo[r]) bad = true;                     // sample precondition of the fast t-statistic failed (checkpoint kernel)
    t.redo[r] = bad ? 1u : 0u;
    if (!bad) {
        const uint32_t m = prefix;                 // peaks; scrappie events n = m + 1; r.events (all means > 0) = m
        if (m == 0) { v.et_n[r] = 0; v.n_events[r] = 0; status = DNB_READ_UNDEFINED; }
        else {
            v.et_n[r] = m + 1; v.n_events[r] = m;
            if (m > cap) status = DNB_READ_OVERFLOW;
        }
        const uint64_t ql = v.q_off[r + 1] - v.q_off[r], rl = v.r_off[r + 1] - v.r_off[r];
        if (ql <= DNB_K || rl < DNB_K) status = DNB_READ_UNDEFINED;
        v.status[r] = status;
    }
}

// ---- K4: events from peaks, one warp per tile ---------------------------------------------------------------------
// Peak k (global index within the read) ends scrappie event k and starts event k+1 (event_detection.c:257-263).
// With every event mean > 0 the filter of event_handling.cpp:549-575 is the identity shifted by one:
// r.events[0] = {0.0, raw from 0}, r.events[k] = {mean_k, raw from peak k-1}, the last scrappie event is dropped.
__global__ void __launch_bounds__(128) seg_events_kernel(DnbBatchView v, DnbSegTiles t) {
    const uint32_t g = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= t.n_tiles) return;
    const uint32_t r = t.tile_read[g];
    if (t.redo[r] || v.status[r] == DNB_READ_OVERFLOW) return;
    const uint32_t c = t.pk_count[g];
    if (c == 0) return;
    const uint32_t prefix = t.tile_prefix[g];
    const uint32_t m = v.n_events[r];
    const uint32_t *pp = t.pk_pos + (size_t)g * DNB_SEG_PEAK_CAP;
    const double *ps = t.pk_sum + (size_t)g * DNB_SEG_PEAK_CAP;
    uint32_t *ev_start = v.ev_start + v.ev_off[r] + r;
    float *ev_mean = v.ev_mean + v.ev_off[r];
    for (uint32_t j = lane; j < c; j += 32) {
        const uint32_t k = prefix + j;
        const uint32_t pos = pp[j];
        const double sm = ps[j];
        const uint32_t ppos = j > 0 ? pp[j - 1] : t.tile_prev_pos[g];
        const double psum = j > 0 ? ps[j - 1] : t.tile_prev_sum[g];
        const unsigned long long span = (unsigned long long)((long long)pos - (long long)ppos);
        const float mean = fDiv(d2f(dSub(sm, psum)), (float)span);     // event_detection.c:225-226
        bool redo = false;
        if (k >= 1 && !((double)mean > 0.0)) redo = true;              // the filter would drop it: serial path
        ev_mean[k] = k >= 1 ? mean : 0.0f;                              // quirk Q1
        if (k == 0) ev_start[0] = 0;
        ev_start[k + 1] = pos;
        if (k == m - 1) {                                               // last scrappie event [pos, N)
            const uint32_t N = v.n_samples[r];
            const float lmean = fDiv(d2f(dSub(t.tot_sum[r], sm)), (float)(unsigned long long)(N - pos));
            if (!((double)lmean > 0.0)) redo = true;
        }
        if (redo) t.redo[r] = 1u;
    }
}

}  // namespace

void dnb_launch_segmentation_serial(const DnbBatchView &v, DnbDetector det, const uint32_t *only_flagged, cudaStream_t s) {
    if (v.n_reads == 0) return;
    dim3 grid((v.n_reads + SEG_THREADS - 1) / SEG_THREADS);
    if (v.raw_i16)
        seg_serial_kernel<true><<<grid, SEG_THREADS, 0, s>>>(v, det, only_flagged);
    else
        seg_serial_kernel<false><<<grid, SEG_THREADS, 0, s>>>(v, det, only_flagged);
}

// Markstein's correction returns the correctly rounded x/w when RN(1/w) is within 2^-54 (relative) of 1/w
static bool fast_div_ok(uint32_t w) {
    if (w < 2 || w > 7) return false;
    const double wd = (double)w, r = 1.0 / wd;
    return std::fabs(std::fma(-wd, r, 1.0)) <= 0x1p-54;
}

// the checkpoints come from the streaming warp scan of seg_scan.cu; DNB_SEG_PARITY_SCAN=0 selects the per-sample
// one-lane-per-read chain (seg_checkpoint_kernel above), kept as the cross-check
bool dnb_seg_parity_scan_enabled(void) {
    static const bool on = !(getenv("DNB_SEG_PARITY_SCAN") != nullptr && getenv("DNB_SEG_PARITY_SCAN")[0] == '0');
    return on;
}

void dnb_launch_segmentation_tiled(const DnbBatchView &v, DnbDetector det, const DnbSegTiles &t, cudaStream_t s,
                                   cudaEvent_t after_checkpoint, cudaEvent_t after_tiles) {
    if (v.n_reads == 0) return;
    const unsigned gr = (v.n_reads + 127) / 128;
    const unsigned gt = (t.n_tiles + SEG_THREADS - 1) / SEG_THREADS;
    const bool fast = fast_div_ok(det.w1) && fast_div_ok(det.w2);
    const bool scanned = dnb_seg_parity_scan_enabled() && dnb_launch_seg_parity_scan(v, t, s) == cudaSuccess;
    if (v.raw_i16) {
        if (!scanned) seg_checkpoint_kernel<true><<<gr, 128, 0, s>>>(v, t);
        if (after_checkpoint) cudaEventRecord(after_checkpoint, s);
        if (fast) seg_tile_kernel<true, true><<<gt, SEG_THREADS, 0, s>>>(v, det, t);
        else seg_tile_kernel<true, false><<<gt, SEG_THREADS, 0, s>>>(v, det, t);
    } else {
        if (!scanned) seg_checkpoint_kernel<false><<<gr, 128, 0, s>>>(v, t);
        if (after_checkpoint) cudaEventRecord(after_checkpoint, s);
        if (fast) seg_tile_kernel<false, true><<<gt, SEG_THREADS, 0, s>>>(v, det, t);
        else seg_tile_kernel<false, false><<<gt, SEG_THREADS, 0, s>>>(v, det, t);
    }
    if (after_tiles) cudaEventRecord(after_tiles, s);
    seg_stitch_kernel<<<(v.n_reads + 3) / 4, 128, 0, s>>>(v, t);
    seg_events_kernel<<<(t.n_tiles + 3) / 4, 128, 0, s>>>(v, t);
    dnb_launch_segmentation_serial(v, det, t.redo, s);
}
