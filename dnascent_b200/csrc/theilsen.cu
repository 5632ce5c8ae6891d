// theilsen.cu -- Theil-Sen refinement of (shift, scale): exact median of all pairwise slopes, one CTA per read.
//
// Replaces estimateScaling_theilSen, /root/reference/src/event_handling.cpp:24-110.
// The reference materialises <= 499 500 slopes and std::sorts them (4 MB, 58 ms per read).  Here the <= 1000
// (x, y) points sit in shared memory and nothing is spilled to HBM.  The median is an exact order statistic, found
// by narrowing the order-preserving 64-bit image of the slopes 12 bits at a time:
//   level 0   histogram of the top 12 bits (sign + exponent) over all pairs        -> the bin holding rank ns/2
//   level 1   histogram of the next 12 bits, over the pairs inside that bin         -> a bin of a few hundred slopes
//   collect   the slopes of that bin are gathered into shared memory and the wanted rank is picked by counting
// Further 12-bit levels are run only while the bin still holds more slopes than the candidate buffer (ties by the
// thousand).  That general search (three sweeps of IEEE divisions) is the fallback: normally select_slope_windowed
// finds the answer in two sweeps over APPROXIMATE quotients (7 instructions each instead of ~35) and certifies it with
// exact divisions of the few hundred candidates around it -- the answer is still the exact order statistic.  Pairs are enumerated circulantly
// (round d pairs point i with point i+d), which keeps the shared-memory reads of a warp on consecutive addresses.
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define TS_THREADS 256
#define TS_MAXP 1000
#define TS_BINS 4096            // 12-bit digits
#define TS_CAND 2048            // candidate buffer (64-bit keys); shares storage with the histogram
#define FULL 0xffffffffu

namespace {

// total order on doubles: -inf < ... < -0 < +0 < ... < +inf < NaN.
// A NaN slope is 0/0: two cleaned points with identical signal and identical model level (about one read in 2000 has
// such a pair).  The reference std::sorts the slopes with operator<, for which NaN is "not less than" anything and
// nothing is less than it; on the reads where it occurs the reference's median is the one obtained with the NaN
// sorted LAST (found by the 2000-read statistical run, profiles/r2n_ea_statistical_parity.json; pinned by
// tests/golden/read_theilsen_nan_slope.npz).  The device's 0/0 is the NEGATIVE canonical NaN, whose raw bit pattern
// would sort first and shift the median by one rank, so NaN gets the largest key explicitly.
__device__ __forceinline__ unsigned long long order_key(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    if (d != d) return ~0ull;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void hist_add(uint32_t *hist, uint32_t d, bool active) {
    // warp-aggregated increment for the digit of the first active lane (at level 0 every slope shares it), plain
    // shared-memory atomics for the rest
    const unsigned act = __ballot_sync(FULL, active);
    if (act == 0) return;
    const int leader = __ffs(act) - 1;
    const uint32_t d0 = __shfl_sync(FULL, d, leader);
    const unsigned same = __ballot_sync(FULL, active && d == d0);
    if ((threadIdx.x & 31) == leader) atomicAdd(&hist[d0], (uint32_t)__popc(same));
    if (active && d != d0) atomicAdd(&hist[d], 1u);
}

struct TsShared {
    double x[TS_MAXP], y[TS_MAXP];
    union {
        uint32_t hist[TS_BINS];
        unsigned long long cand[TS_CAND];
    } u;
    uint32_t hist8[256];         // digit histogram of select_among_candidates
    unsigned long long sel_prefix;
    uint32_t sel_rank;
    unsigned long long prefix;   // key bits decided so far
    uint32_t rank;               // wanted rank inside the current bin
    uint32_t count;              // slopes inside the current bin
    uint32_t n_cand;
    unsigned long long answer;
};

// dy/dx within TS_APPROX_ULPS ulps of the IEEE quotient, in 7 instructions instead of the ~35 of the division
// subroutine: reciprocal seed (2^-20) + two Newton steps (-> ~1 ulp) + one multiply.  Where the bound cannot be
// vouched for (dx == 0, non-finite or subnormal-range results) the exact quotient is returned instead.
#define TS_APPROX_ULPS 4096ull
__device__ __forceinline__ double slope_approx(double dy, double dx) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dx));
    double e = __fma_rn(-dx, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-dx, r, 1.0);
    r = __fma_rn(r, e, r);
    const double q = dMul(dy, r);
    const double aq = fabs(q);
    if (!(aq < 1.0e290) || (aq < 1.0e-290 && dy != 0.0)) return dDiv(dy, dx);
    return q;
}

// visits every unordered pair once; f(key) is called by all lanes of a warp together (ok == false for padding lanes).
// kApprox: keys of slope_approx (within TS_APPROX_ULPS of the exact key) instead of the IEEE quotient's
template <bool kApprox = false, class F>
__device__ __forceinline__ void for_each_slope(const TsShared &sm, uint32_t np, F f) {
    const uint32_t full_rounds = (np - 1) / 2;
    const uint32_t lane_base = threadIdx.x;
    for (uint32_t d = 1; d <= full_rounds; d++) {
        for (uint32_t base = 0; base < np; base += TS_THREADS) {
            const uint32_t i = base + lane_base;
            const bool ok = i < np;
            unsigned long long key = 0;
            if (ok) {
                uint32_t j = i + d;
                if (j >= np) j -= np;
                // oriented lower index first, exactly as the reference's nested loop (:67-75): dx, dy and the signed
                // zero of dy/dx are the reference's
                const uint32_t lo = min(i, j), hi = max(i, j);
                const double dy = dSub(sm.y[lo], sm.y[hi]), dx = dSub(sm.x[lo], sm.x[hi]);
                key = order_key(kApprox ? slope_approx(dy, dx) : dDiv(dy, dx));
            }
            f(key, ok, i, d);
        }
    }
    if ((np & 1u) == 0) {   // even np: the half round d = np/2 pairs i < np/2 with i + np/2
        const uint32_t d = np / 2;
        for (uint32_t base = 0; base < d; base += TS_THREADS) {
            const uint32_t i = base + lane_base;
            const bool ok = i < d;
            unsigned long long key = 0;
            if (ok) {
                const double dy = dSub(sm.y[i], sm.y[i + d]), dx = dSub(sm.x[i], sm.x[i + d]);
                key = order_key(kApprox ? slope_approx(dy, dx) : dDiv(dy, dx));
            }
            f(key, ok, i, d);
        }
    }
}
// the exact key of the pair (i, i + d mod np) that for_each_slope visited
__device__ __forceinline__ unsigned long long exact_key(const TsShared &sm, uint32_t np, uint32_t i, uint32_t d) {
    uint32_t j = i + d;
    if (j >= np) j -= np;
    const uint32_t lo = min(i, j), hi = max(i, j);
    return order_key(dDiv(dSub(sm.y[lo], sm.y[hi]), dSub(sm.x[lo], sm.x[hi])));
}

// rank-th smallest (0-based) of the n keys in sm.u.cand: radix select, eight 8-bit digits from the top (n * 8 key
// visits; the first version ranked every candidate against every other one, n^2 compares, which ncu showed to be ~40 %
// of the kernel's instructions)
__device__ __forceinline__ void select_among_candidates(TsShared &sm, uint32_t n, uint32_t rank) {
    const int tid = threadIdx.x;
    if (tid == 0) { sm.sel_prefix = 0ull; sm.sel_rank = rank; }
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        sm.hist8[tid] = 0;                                     // TS_THREADS == 256
        __syncthreads();
        const unsigned long long prefix = sm.sel_prefix;
        for (uint32_t c0 = 0; c0 < n; c0 += TS_THREADS) {
            const uint32_t c = c0 + tid;
            const unsigned long long k = c < n ? sm.u.cand[c] : 0ull;
            hist_add(sm.hist8, (uint32_t)(k >> shift) & 0xFFu, c < n && (k & himask) == prefix);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) sum += sm.hist8[tid * 8 + q];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += t;
            }
            const uint32_t excl = incl - sum, want = sm.sel_rank;
            if (want >= excl && want < incl) {
                uint32_t rem = want - excl, q = 0;
                for (; q < 7; q++) {
                    const uint32_t c = sm.hist8[tid * 8 + q];
                    if (rem < c) break;
                    rem -= c;
                }
                sm.sel_prefix = prefix | ((unsigned long long)(tid * 8 + q) << shift);
                sm.sel_rank = rem;
            }
        }
        __syncthreads();
    }
    if (tid == 0) sm.answer = sm.sel_prefix;
    __syncthreads();
}

// Windowed shortcut: two sweeps over APPROXIMATE slopes, exact divisions only for the few hundred candidates.
// The median of one circulant round of (exact) slopes places a window of 2^53 keys -- about one binade either side
// of it -- whose 4096 linear bins are histogrammed over the approximate keys; the bin [lo, hi) holding the wanted rank
// is widened by the approximation's error bound, the pairs whose approximate key falls in [lo - delta, hi + delta) get
// their exact quotient computed and are ranked exactly, the pairs approximately below lo - delta are counted.  A pair
// approximately below lo - delta is exactly below lo, one approximately at or above hi + delta is exactly at or above
// hi, so if the candidate picked by exact rank lies in [lo, hi) its global rank is exactly the wanted one.  Otherwise
// (answer in the fuzzy margin, window missed, candidate buffer too small) the caller runs the general exact search.
__device__ bool select_slope_windowed(TsShared &sm, uint32_t np, uint32_t kth, unsigned long long *out) {
    const int tid = threadIdx.x;
    const uint32_t d = (np - 1) / 2;                  // a full round: every point paired with the one d places on
    for (uint32_t i = tid; i < np; i += TS_THREADS) sm.u.cand[i] = exact_key(sm, np, i, d);
    __syncthreads();
    select_among_candidates(sm, np, np / 2);
    const unsigned long long centre = sm.answer;
    const unsigned long long half = 1ull << 52;
    const unsigned long long k0 = centre < half ? 0ull : (centre > ~0ull - half ? ~0ull - 2 * half + 1 : centre - half);
    __syncthreads();
    for (int i = tid; i < TS_BINS; i += TS_THREADS) sm.u.hist[i] = 0;
    if (tid == 0) { sm.count = 0; sm.n_cand = 0; }
    __syncthreads();
    uint32_t below = 0;
    for_each_slope<true>(sm, np, [&](unsigned long long key, bool ok, uint32_t, uint32_t) {
        const unsigned long long off = key - k0;
        below += (ok && key < k0) ? 1u : 0u;
        if (ok && key >= k0 && (off >> 53) == 0) atomicAdd(&sm.u.hist[(uint32_t)(off >> 41)], 1u);
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(FULL, below, o);
    if ((tid & 31) == 0) atomicAdd(&sm.count, below);
    __syncthreads();
    below = sm.count;
    __syncthreads();
    if (tid < 32) {
        const uint32_t per = TS_BINS / 32;
        uint32_t sum = 0;
        for (uint32_t q = 0; q < per; q++) sum += sm.u.hist[tid * per + q];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (tid >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const bool inside = kth >= below && kth - below < total;
        if (tid == 0) sm.rank = 0xffffffffu;
        __syncwarp();
        if (inside) {
            const uint32_t excl = incl - sum, rank = kth - below;
            if (rank >= excl && rank < incl) {
                uint32_t rem = rank - excl, q = 0;
                for (; q < per - 1; q++) {
                    const uint32_t c = sm.u.hist[tid * per + q];
                    if (rem < c) break;
                    rem -= c;
                }
                const uint32_t bin = tid * per + q;
                sm.prefix = bin;
                sm.rank = rem;
                sm.count = sm.u.hist[bin];
            }
        }
    }
    __syncthreads();
    // the first and the last bin have no room for the margin inside the window
    if (sm.rank == 0xffffffffu || sm.count + 64 > TS_CAND || sm.prefix == 0 || sm.prefix == TS_BINS - 1) return false;
    const uint32_t bin = (uint32_t)sm.prefix;
    const unsigned long long lo = k0 + ((unsigned long long)bin << 41), hi = lo + (1ull << 41);
    const unsigned long long delta = TS_APPROX_ULPS;
    __syncthreads();
    if (tid == 0) { sm.count = 0; sm.n_cand = 0; }
    __syncthreads();
    uint32_t n_below = 0;
    for_each_slope<true>(sm, np, [&](unsigned long long key, bool ok, uint32_t i, uint32_t dd) {
        if (!ok) return;
        if (key < lo - delta) { n_below++; return; }
        if (key < hi + delta) {
            const uint32_t slot = atomicAdd(&sm.n_cand, 1u);
            if (slot < TS_CAND) sm.u.cand[slot] = exact_key(sm, np, i, dd);
        }
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_below += __shfl_xor_sync(FULL, n_below, o);
    if ((tid & 31) == 0) atomicAdd(&sm.count, n_below);
    __syncthreads();
    n_below = sm.count;
    const uint32_t n_cand = sm.n_cand;
    __syncthreads();
    if (n_cand > TS_CAND || kth < n_below || kth - n_below >= n_cand) return false;
    select_among_candidates(sm, n_cand, kth - n_below);
    const unsigned long long ans = sm.answer;
    __syncthreads();
    if (ans < lo || ans >= hi) return false;          // in the margin: its global rank is not certified
    *out = ans;
    return true;
}

// exact k-th smallest slope over all pairs
__device__ unsigned long long select_slope(TsShared &sm, uint32_t np, uint32_t kth, int mode) {
    const int tid = threadIdx.x;
    if (mode == 0) {
        unsigned long long fast;
        if (select_slope_windowed(sm, np, kth, &fast)) return fast;
        __syncthreads();
    }
    if (tid == 0) { sm.prefix = 0ull; sm.rank = kth; sm.count = np * (np - 1) / 2; }
    __syncthreads();
    for (int level = 0; level < 6; level++) {
        const int shift = level < 5 ? 52 - 12 * level : 0;            // digits: bits 63-52, 51-40, 39-28, 27-16, 15-4, 3-0
        const uint32_t dmask = level < 5 ? 0xFFFu : 0xFu;
        const unsigned long long himask = level == 0 ? 0ull : (~0ull << (level < 5 ? shift + 12 : 4));
        const unsigned long long prefix = sm.prefix;
        if (level > 0 && sm.count <= TS_CAND) {
            // few enough slopes share the decided bits: gather them and finish
            if (tid == 0) sm.n_cand = 0;
            __syncthreads();
            for_each_slope(sm, np, [&](unsigned long long key, bool ok, uint32_t, uint32_t) {
                if (ok && (key & himask) == prefix) sm.u.cand[atomicAdd(&sm.n_cand, 1u)] = key;
            });
            __syncthreads();
            select_among_candidates(sm, sm.n_cand, sm.rank);
            return sm.answer;
        }
        for (int i = tid; i < TS_BINS; i += TS_THREADS) sm.u.hist[i] = 0;
        __syncthreads();
        for_each_slope(sm, np, [&](unsigned long long key, bool ok, uint32_t, uint32_t) {
            hist_add(sm.u.hist, (uint32_t)(key >> shift) & dmask, ok && (key & himask) == prefix);
        });
        __syncthreads();
        if (tid < 32) {
            // warp 0 finds the bin holding the wanted rank: per-lane partial sums over 128-bin stripes, then a scan
            const uint32_t per = TS_BINS / 32;
            uint32_t sum = 0;
            for (uint32_t q = 0; q < per; q++) sum += sm.u.hist[tid * per + q];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += t;
            }
            const uint32_t excl = incl - sum, rank = sm.rank;
            const bool mine = rank >= excl && rank < incl;
            if (mine) {
                uint32_t rem = rank - excl, q = 0;
                for (; q < per - 1; q++) {
                    const uint32_t c = sm.u.hist[tid * per + q];
                    if (rem < c) break;
                    rem -= c;
                }
                const uint32_t digit = tid * per + q;
                sm.prefix = prefix | ((unsigned long long)digit << shift);
                sm.rank = rem;
                sm.count = sm.u.hist[digit];
            }
        }
        __syncthreads();
    }
    return sm.prefix;   // all 64 bits decided: the bin holds copies of one value
}

__global__ void __launch_bounds__(TS_THREADS) theil_sen_kernel(DnbBatchView v, DnbModelDev m, DnbTsArgs a) {
    __shared__ TsShared sm;
    const uint32_t r = v.order[blockIdx.x];
    const int tid = threadIdx.x;
    const int st = v.status[r];
    if (st == DNB_READ_UNDEFINED || st == DNB_READ_OVERFLOW) return;
    const double shift = a.rough_shift[r], scale = a.rough_scale[r];
    const uint32_t n = a.n_cleaned[r];
    const uint32_t maxPoints = TS_MAXP, trim = 50;
    if (n < maxPoints) {                                   // :33 short reads keep the rough scaling
        if (tid == 0) { a.shift[r] = shift; a.scale[r] = scale; }
        return;
    }
    const double *sig = a.cl_signal + a.cl_off[r];
    const uint32_t *rk = a.cl_rank + a.cl_off[r];
    const uint32_t eff = n - 2 * trim;
    uint32_t skip = 1, np = eff;
    if (eff > maxPoints) { skip = eff / maxPoints; np = maxPoints; }
    for (uint32_t j = tid; j < np; j += TS_THREADS) {
        const uint32_t i = trim + j * skip;
        sm.x[j] = dDiv(dSub(sig[i], shift), scale);        // :51
        sm.y[j] = m.mean[rk[i]];                            // :58
    }
    __syncthreads();

    // median slope: element ns/2 of the ascending sort of dy/dx over all i<j (:67-78)
    const uint32_t ns = np * (np - 1) / 2;
    const double slope = key_to_double(select_slope(sm, np, ns / 2, a.mode));
    __syncthreads();

    // median intercept: element np/2 of y - slope*x (:81-87); np <= TS_MAXP <= TS_CAND
    for (uint32_t i = tid; i < np; i += TS_THREADS)
        sm.u.cand[i] = order_key(dSub(sm.y[i], dMul(slope, sm.x[i])));   // :83, not fused
    __syncthreads();
    select_among_candidates(sm, np, np / 2);
    const double icpt = key_to_double(sm.answer);

    if (tid == 0) {
        double o_shift, o_scale;
        if (slope == 0.) {                                  // :90-95
            o_shift = -1.; o_scale = -1.;
        } else {
            const double scale_corr = dDiv(1., slope);
            const double shift_corr = dDiv(-icpt, slope);
            o_shift = dAdd(shift, dMul(shift_corr, scale));
            o_scale = dMul(scale, scale_corr);
        }
        a.shift[r] = o_shift;
        a.scale[r] = o_scale;
        if (o_shift == -1. && st == DNB_READ_OK) v.status[r] = DNB_READ_SCALE_FAIL;   // event_handling.cpp:604
    }
}

}  // namespace

void dnb_launch_theil_sen(const DnbBatchView &v, const DnbModelDev &m, const DnbTsArgs &a, cudaStream_t s) {
    if (v.n_reads == 0) return;
    theil_sen_kernel<<<v.n_reads, TS_THREADS, 0, s>>>(v, m, a);
}
