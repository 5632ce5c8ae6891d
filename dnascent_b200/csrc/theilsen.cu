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
#define NSP_FULL 256            // the NaN path sorts at most this many slopes literally (one thread)
#include "nan_sort_path.cuh"

#define TS_THREADS 256
#define TS_MAXP 1000
#define TS_BINS 4096            // 12-bit digits
#define TS_CAND 2048            // candidate buffer (64-bit keys); shares storage with the histogram
#define FULL 0xffffffffu

namespace {

// total order on doubles: -inf < ... < -0 < +0 < ... < +inf < NaN.
// A NaN slope is 0/0: two cleaned points with identical signal and identical model level (about one read in 1000 has
// such a pair).  The reference std::sorts the slopes with operator<, which is outside std::sort's contract with a NaN
// in the range: where the NaN ends up -- and with it whether the median is the rank ns/2 or ns/2 - 1 non-NaN slope --
// is whatever libstdc++'s introsort does with that particular sequence of slopes (found by the 2000-read statistical
// run: after the median on read 1797, before it on read 1803; tests/golden/read_theilsen_nan_slope{,_b}.npz).  Here
// NaN gets the largest key, i.e. this kernel computes the "NaN last" answer, flags the read, and
// theil_sen_nan_follow_kernel below re-derives the median by following the NaN through the introsort.
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
    uint32_t n_nan;              // 0/0 slopes seen (any nonzero value sends the read to theil_sen_nan_kernel)
};

// dy/dx within TS_APPROX_ULPS ulps of the IEEE quotient, in 7 instructions instead of the ~35 of the division
// subroutine: reciprocal seed (2^-20) + two Newton steps (-> ~1 ulp) + one multiply.  Where the bound cannot be
// vouched for (dx == 0, non-finite or subnormal-range results) the exact quotient is returned instead.
#define TS_APPROX_ULPS 4096ull
__device__ __forceinline__ double slope_approx(double dy, double dx, uint32_t *n_nan) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dx));
    double e = __fma_rn(-dx, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-dx, r, 1.0);
    r = __fma_rn(r, e, r);
    const double q = dMul(dy, r);
    const double aq = fabs(q);
    if (!(aq < 1.0e290) || (aq < 1.0e-290 && dy != 0.0)) {
        const double exact = dDiv(dy, dx);
        if (exact != exact) atomicAdd(n_nan, 1u);       // 0/0: on this rare branch only, the common path pays nothing
        return exact;
    }
    return q;
}

// visits every unordered pair once; f(key) is called by all lanes of a warp together (ok == false for padding lanes).
// kApprox: keys of slope_approx (within TS_APPROX_ULPS of the exact key) instead of the IEEE quotient's
template <bool kApprox = false, class F>
__device__ __forceinline__ void for_each_slope(TsShared &sm, uint32_t np, F f) {
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
                key = order_key(kApprox ? slope_approx(dy, dx, &sm.n_nan) : dDiv(dy, dx));
                if (!kApprox && key == ~0ull) atomicAdd(&sm.n_nan, 1u);
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
                key = order_key(kApprox ? slope_approx(dy, dx, &sm.n_nan) : dDiv(dy, dx));
                if (!kApprox && key == ~0ull) atomicAdd(&sm.n_nan, 1u);
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

// the <= 1000 (x, y) points of a read (event_handling.cpp:35-64); 0 if the read keeps its rough scaling (:33)
__device__ __forceinline__ uint32_t ts_load_points(TsShared &sm, const DnbModelDev &m, const DnbTsArgs &a, uint32_t r, double shift, double scale) {
    const uint32_t n = a.n_cleaned[r];
    const uint32_t maxPoints = TS_MAXP, trim = 50;
    if (n < maxPoints) return 0;
    const double *sig = a.cl_signal + a.cl_off[r];
    const uint32_t *rk = a.cl_rank + a.cl_off[r];
    const uint32_t eff = n - 2 * trim;
    uint32_t skip = 1, np = eff;
    if (eff > maxPoints) { skip = eff / maxPoints; np = maxPoints; }
    for (uint32_t j = threadIdx.x; j < np; j += TS_THREADS) {
        const uint32_t i = trim + j * skip;
        sm.x[j] = dDiv(dSub(sig[i], shift), scale);        // :51
        sm.y[j] = m.mean[rk[i]];                            // :58
    }
    __syncthreads();
    return np;
}

// median intercept for the given median slope and the refined (shift, scale) (event_handling.cpp:80-108); block-wide.
// `st` is the read's status before Theil-Sen.
__device__ __forceinline__ void ts_finish(TsShared &sm, const DnbBatchView &v, const DnbTsArgs &a, uint32_t r, uint32_t np, double slope,
                                          double shift, double scale, int st) {
    const int tid = threadIdx.x;
    // element np/2 of y - slope*x (:81-87); np <= TS_MAXP <= TS_CAND
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
        if (st == DNB_READ_OK) v.status[r] = (o_shift == -1.) ? DNB_READ_SCALE_FAIL : DNB_READ_OK;   // event_handling.cpp:604
    }
}

__global__ void __launch_bounds__(TS_THREADS) theil_sen_kernel(DnbBatchView v, DnbModelDev m, DnbTsArgs a) {
    __shared__ TsShared sm;
    const uint32_t r = v.order[blockIdx.x];
    const int tid = threadIdx.x;
    const int st = v.status[r];
    if (st == DNB_READ_UNDEFINED || st == DNB_READ_OVERFLOW) return;
    const double shift = a.rough_shift[r], scale = a.rough_scale[r];
    if (tid == 0) sm.n_nan = 0;
    const uint32_t np = ts_load_points(sm, m, a, r, shift, scale);
    if (np == 0) {                                         // :33 short reads keep the rough scaling
        if (tid == 0) { a.shift[r] = shift; a.scale[r] = scale; }
        return;
    }
    // median slope: element ns/2 of the ascending sort of dy/dx over all i<j (:67-78)
    const uint32_t ns = np * (np - 1) / 2;
    const double slope = key_to_double(select_slope(sm, np, ns / 2, a.mode));
    __syncthreads();
    // a 0/0 slope was seen: the answer above is the "NaN sorted last" one.  theil_sen_nan_follow_kernel decides whether
    // it stands; the other possible answer (NaN in front of the median: rank ns/2 - 1 of the non-NaN slopes) is
    // selected here, where the device is full of other CTAs, rather than by a lone CTA later
    const bool flagged = sm.n_nan != 0 && a.nan_list != nullptr;      // uniform: read after the barrier
    double slope_alt = 0.0;
    if (flagged) {
        slope_alt = key_to_double(select_slope(sm, np, ns / 2 - 1, a.mode));
        __syncthreads();
    }
    ts_finish(sm, v, a, r, np, slope, shift, scale, st);
    if (tid == 0 && flagged) {
        const uint32_t slot = atomicAdd(a.nan_count, 1u);
        if (slot < a.nan_cap) {
            a.nan_list[4 * slot] = r; a.nan_list[4 * slot + 1] = (uint32_t)st; a.nan_list[4 * slot + 2] = 0u;
            reinterpret_cast<double *>(a.nan_list + 4 * a.nan_cap)[a.nan_cap + slot] = slope_alt;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Reads with a NaN slope (see order_key).  theil_sen_nan_follow_kernel materialises the <= 499 500 slopes in the
// reference's push order (:67-75) in a scratch slot and follows the NaN through libstdc++'s introsort
// (nan_sort_path.cuh): only the partitions of the ranges that contain the NaN are applied, each as prefix counts +
// parallel swaps by the CTA -- nsp_partition_lists is the serial statement of what is done here -- and the last
// <= NSP_FULL elements are sorted literally by one thread.  That tells where element ns/2 of the array std::sort would
// have produced comes from: the literally sorted range (read off), left of it (the NaN is behind the median: the
// "NaN last" answer stands) or right of it (the NaN is in front: the median is the rank ns/2 - 1 non-NaN slope).
// theil_sen_nan_finish_kernel then redoes the intercept and the refined scalings of the reads whose answer changes
// (the rank ns/2 - 1 slope was selected by theil_sen_kernel when it flagged the read).
// One CTA per flagged read, grid-strided over the list; both kernels are launched after every theil_sen_kernel and
// normally find a list of a few reads (about one read in 1000 is flagged).
// Not emulated (the "NaN last" answer stands): more than one NaN, the NaN chosen as the pivot of a range of more than
// NSP_PIVOT_MAX slopes, the introsort depth limit reached on the NaN's branch.
// ---------------------------------------------------------------------------------------------------------------
#define NAN_THREADS 1024
#define NAN_PER 4               // elements per thread and tile (loads in flight)
struct NanShared {
    double x[TS_MAXP], y[TS_MAXP];
    long first, last, nan_pos, nan_first;
    int depth, state;            // state: 0 partition next, 1 done (range sorted literally), 2 give up
    uint32_t cnt[NAN_PER * NAN_THREADS / 32];       // per (sub-tile, warp): stops of the left pointer | stops of the right pointer << 16
    uint32_t base_a, base_d, n_nan;
    double buf[NSP_FULL + NSP_THRESHOLD];
};

__global__ void __launch_bounds__(NAN_THREADS) theil_sen_nan_follow_kernel(DnbBatchView v, DnbModelDev m, DnbTsArgs a) {
    __shared__ NanShared sh;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t n_list = min(*a.nan_count, a.nan_cap);
    double *sl = a.nan_scratch + (size_t)blockIdx.x * DNB_TS_NAN_SLOT_DOUBLES;
    int *listA = reinterpret_cast<int *>(sl + DNB_TS_MAX_SLOPES);
    int *listD = listA + DNB_TS_MAX_SLOPES;
    double *nan_val = reinterpret_cast<double *>(a.nan_list + 4 * a.nan_cap);
    for (uint32_t e = blockIdx.x; e < n_list; e += gridDim.x) {
        const uint32_t r = a.nan_list[4 * e];
        const double shift = a.rough_shift[r], scale = a.rough_scale[r];
        __syncthreads();
        if (tid == 0) { sh.n_nan = 0; sh.nan_first = -1; a.nan_list[4 * e + 2] = 0u; }
        // the points (as ts_load_points; this CTA is wider than TS_THREADS)
        const uint32_t nc = a.n_cleaned[r];
        uint32_t np = 0;
        if (nc >= TS_MAXP) {
            const double *sig = a.cl_signal + a.cl_off[r];
            const uint32_t *rk = a.cl_rank + a.cl_off[r];
            const uint32_t eff = nc - 100;
            uint32_t skip = 1;
            np = eff;
            if (eff > TS_MAXP) { skip = eff / TS_MAXP; np = TS_MAXP; }
            for (uint32_t j = tid; j < np; j += NAN_THREADS) {
                const uint32_t i = 50 + j * skip;
                sh.x[j] = dDiv(dSub(sig[i], shift), scale);
                sh.y[j] = m.mean[rk[i]];
            }
        }
        __syncthreads();
        if (np < 2) continue;
        const long n = (long)np * (np - 1) / 2;
        // slopes in push order: pair (i, j), i < j, at index i*np - i*(i+1)/2 + (j - i - 1)
        for (uint32_t i = wid; i + 1 < np; i += NAN_THREADS / 32) {        // a warp per row
            const long row = (long)i * np - (long)i * (i + 1) / 2 - (i + 1);
            const double xi = sh.x[i], yi = sh.y[i];
            for (uint32_t j = i + 1 + lane; j < np; j += 32) {
                const double q = dDiv(dSub(yi, sh.y[j]), dSub(xi, sh.x[j]));
                sl[row + j] = q;
                if (q != q) { atomicAdd(&sh.n_nan, 1u); sh.nan_first = row + j; }
            }
        }
        __syncthreads();
        if (sh.n_nan != 1) continue;                       // uniform: shared value read after the barrier
        if (tid == 0) { sh.first = 0; sh.last = n; sh.depth = nsp_lg(n) * 2; sh.nan_pos = sh.nan_first; sh.state = 0; }
        __syncthreads();
        while (true) {
            // ---- one level of std::__introsort_loop on the range holding the NaN ----
            if (sh.last - sh.first <= NSP_FULL) {
                // small enough: the rest of the introsort and the insertion pass literally, by one thread, on a copy in
                // shared memory (on global memory its ~3000 dependent accesses were most of this kernel's time).  The
                // insertion pass is guarded against v[0] for indices below 16, so a range that starts there is
                // staged from index 0.
                const long first = sh.first, last = sh.last;
                const long s0 = first < NSP_THRESHOLD ? 0 : first;
                for (long i = s0 + tid; i < last; i += NAN_THREADS) sh.buf[i - s0] = sl[i];
                __syncthreads();
                if (tid == 0) {
                    double *vv = sh.buf - s0;
                    bool ok = nsp_introsort_full(vv, first, last, sh.depth);
                    ok = ok && nsp_insertion_pass(vv, first, last);
                    sh.state = ok ? 1 : 2;
                }
                __syncthreads();
                for (long i = s0 + tid; i < last; i += NAN_THREADS) sl[i] = sh.buf[i - s0];
                __syncthreads();
                break;
            }
            if (tid == 0) {
                const long first = sh.first, last = sh.last;
                if (sh.depth == 0) {
                    sh.state = 2;
                } else {
                    sh.depth--;
                    const long mid = first + (last - first) / 2;
                    nsp_median_to_first(sl, first, first + 1, mid, last - 1);
                    if (sl[first] != sl[first]) {
                        // the NaN is the pivot: the partition orders nothing, the whole range is sorted literally
                        if (last - first > NSP_PIVOT_MAX) sh.state = 2;
                        else {
                            const long cut = nsp_partition(sl, first + 1, last, first);
                            bool ok = nsp_introsort_full(sl, cut, last, sh.depth);
                            ok = nsp_introsort_full(sl, first, cut, sh.depth) && ok;
                            ok = nsp_insertion_pass(sl, first, last) && ok;
                            sh.state = ok ? 1 : 2;
                        }
                    } else {
                        if (sl[first + 1] != sl[first + 1]) sh.nan_pos = first + 1;       // the median-of-3 may have moved it
                        else if (sl[mid] != sl[mid]) sh.nan_pos = mid;
                        else if (sl[last - 1] != sl[last - 1]) sh.nan_pos = last - 1;
                    }
                }
                sh.base_a = 0; sh.base_d = 0;
            }
            __syncthreads();
            if (sh.state != 0) break;
            // ---- std::__unguarded_partition(first + 1, last, first) in closed form (nsp_partition_lists) ----
            const long lo = sh.first + 1, hi = sh.last;
            const double p = sl[sh.first];
            const unsigned below = (1u << lane) - 1u;
            for (long base = lo; base < hi; base += NAN_PER * NAN_THREADS) {
                double x[NAN_PER];
#pragma unroll
                for (int q = 0; q < NAN_PER; q++) {
                    const long i = base + q * NAN_THREADS + tid;
                    x[q] = i < hi ? sl[i] : 0.0;
                }
                unsigned ma[NAN_PER], md[NAN_PER];
#pragma unroll
                for (int q = 0; q < NAN_PER; q++) {
                    const bool ok = base + q * NAN_THREADS + tid < hi;
                    ma[q] = __ballot_sync(FULL, ok && !(x[q] < p));
                    md[q] = __ballot_sync(FULL, ok && !(p < x[q]));
                    if (lane == 0) sh.cnt[q * 32 + wid] = (uint32_t)__popc(ma[q]) | ((uint32_t)__popc(md[q]) << 16);   // a tile holds 4096 < 2^16
                }
                __syncthreads();
                // exclusive prefix of the 4 x 32 (sub-tile, warp) counts: lane l scans warp l's count, every warp redundantly
                uint32_t oa[NAN_PER], od[NAN_PER];
                uint32_t run_a = sh.base_a, run_d = sh.base_d;
#pragma unroll
                for (int q = 0; q < NAN_PER; q++) {
                    const uint32_t c = sh.cnt[q * 32 + lane];
                    uint32_t inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, inc, o);
                        if (lane >= o) inc += t;
                    }
                    const uint32_t mine = __shfl_sync(FULL, inc - c, wid), tot = __shfl_sync(FULL, inc, 31);
                    oa[q] = run_a + (mine & 0xffffu); od[q] = run_d + (mine >> 16);
                    run_a += tot & 0xffffu; run_d += tot >> 16;
                }
#pragma unroll
                for (int q = 0; q < NAN_PER; q++) {
                    const long i = base + q * NAN_THREADS + tid;
                    if ((ma[q] >> lane) & 1u) listA[oa[q] + __popc(ma[q] & below)] = (int)i;
                    if ((md[q] >> lane) & 1u) listD[od[q] + __popc(md[q] & below)] = (int)i;
                }
                __syncthreads();
                if (tid == 0) { sh.base_a = run_a; sh.base_d = run_d; }   // read after the next barrier
            }
            __syncthreads();
            const long ta = sh.base_a, td = sh.base_d;
            // m = the number of k in 1..min(ta, td) with a_k < d_k (monotone): 32 probes per step, one per lane, every warp
            // on the same data (a binary search by dependent global loads cost ~40 us per level)
            long mlo = 0, mhi = ta < td ? ta : td;
            while (mlo < mhi) {
                const long step = (mhi - mlo + 31) / 32;
                long k = mlo + (long)(lane + 1) * step;
                if (k > mhi) k = mhi;
                const int t = __popc(__ballot_sync(FULL, listA[k - 1] < listD[td - k]));      // lanes 0..t-1 hold
                const long nlo = t == 0 ? mlo : min(mlo + (long)t * step, mhi);
                const long nhi = t == 32 ? mhi : min(mlo + (long)(t + 1) * step, mhi) - 1;
                mlo = nlo; mhi = nhi;
            }
            const long mm = mlo;
            for (long k0 = 0; k0 < mm; k0 += NAN_PER * NAN_THREADS) {
                int ia[NAN_PER], id[NAN_PER];
                double va[NAN_PER], vd[NAN_PER];
#pragma unroll
                for (int q = 0; q < NAN_PER; q++) {
                    const long k = k0 + q * NAN_THREADS + tid;
                    ia[q] = k < mm ? listA[k] : -1;
                    id[q] = k < mm ? listD[td - 1 - k] : -1;
                }
#pragma unroll
                for (int q = 0; q < NAN_PER; q++)
                    if (ia[q] >= 0) { va[q] = sl[ia[q]]; vd[q] = sl[id[q]]; }
#pragma unroll
                for (int q = 0; q < NAN_PER; q++)
                    if (ia[q] >= 0) {                            // the pairs are disjoint: no position is in two swaps
                        sl[ia[q]] = vd[q]; sl[id[q]] = va[q];
                        if (va[q] != va[q]) sh.nan_pos = id[q];   // exactly one NaN: at most one thread writes
                        if (vd[q] != vd[q]) sh.nan_pos = ia[q];
                    }
            }
            __syncthreads();
            if (tid == 0) {
                long cut = mm >= 1 ? (long)listD[td - mm] : hi;
                if (mm < ta && listA[mm] < cut) cut = listA[mm];
                if (sh.nan_pos < cut) sh.last = cut; else sh.first = cut;
            }
            __syncthreads();
        }
        if (sh.state != 1) continue;
        if (tid == 0) {
            const long med = n / 2;
            if (med >= sh.last) a.nan_list[4 * e + 2] = 1u;                                   // rank ns/2 - 1 of the non-NaN slopes
            else if (med >= sh.first) { a.nan_list[4 * e + 2] = 2u; nan_val[e] = sl[med]; }   // read off
        }
    }
}

__global__ void __launch_bounds__(TS_THREADS) theil_sen_nan_finish_kernel(DnbBatchView v, DnbModelDev m, DnbTsArgs a) {
    __shared__ TsShared sm;
    const uint32_t n_list = min(*a.nan_count, a.nan_cap);
    const double *nan_val = reinterpret_cast<const double *>(a.nan_list + 4 * a.nan_cap);
    for (uint32_t e = blockIdx.x; e < n_list; e += gridDim.x) {
        const uint32_t action = a.nan_list[4 * e + 2];
        if (action == 0u) continue;
        const uint32_t r = a.nan_list[4 * e];
        const int st = (int)a.nan_list[4 * e + 1];
        const double shift = a.rough_shift[r], scale = a.rough_scale[r];
        __syncthreads();
        if (threadIdx.x == 0) sm.n_nan = 0;
        const uint32_t np = ts_load_points(sm, m, a, r, shift, scale);
        if (np == 0) continue;
        const double slope = action == 2u ? nan_val[e] : nan_val[a.nan_cap + e];   // read off the sorted range / selected by theil_sen_kernel
        ts_finish(sm, v, a, r, np, slope, shift, scale, st);
    }
}

}  // namespace

void dnb_launch_theil_sen(const DnbBatchView &v, const DnbModelDev &m, const DnbTsArgs &a, cudaStream_t s) {
    if (v.n_reads == 0) return;
    if (a.nan_list) cudaMemsetAsync(a.nan_count, 0, sizeof(uint32_t), s);
    theil_sen_kernel<<<v.n_reads, TS_THREADS, 0, s>>>(v, m, a);
    if (a.nan_list) {
        theil_sen_nan_follow_kernel<<<a.nan_slots, NAN_THREADS, 0, s>>>(v, m, a);
        theil_sen_nan_finish_kernel<<<a.nan_slots, TS_THREADS, 0, s>>>(v, m, a);
    }
}
