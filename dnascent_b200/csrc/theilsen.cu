// theilsen.cu -- Theil-Sen refinement of (shift, scale): exact median of all pairwise slopes, one CTA per read.
//
// Replaces estimateScaling_theilSen, /root/reference/src/event_handling.cpp:24-110.
// The reference materialises <= 499 500 slopes and std::sorts them (4 MB, 58 ms per read).  Here the <= 1000
// (x, y) points sit in shared memory and the median is found by an exact 8 x 8-bit radix select over the
// order-preserving 64-bit image of the slopes, recomputing the IEEE divisions in every pass; nothing is
// spilled to HBM.  A pass votes per warp first (in the leading passes every slope shares its digit), so the
// shared-memory histogram sees one atomic per warp instead of 32 on the same address.
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define TS_THREADS 256
#define TS_MAXP 1000
#define FULL 0xffffffffu

namespace {

// total order on doubles: -inf < ... < -0 < +0 < ... < +inf   (NaN never compares in the reference: undefined there)
__device__ __forceinline__ unsigned long long order_key(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void hist_add(uint32_t *hist, uint32_t d, bool active) {
    // warp-aggregated increment for the digit of the first active lane, plain atomics for the rest
    const unsigned act = __ballot_sync(FULL, active);
    if (act == 0) return;
    const int leader = __ffs(act) - 1;
    const uint32_t d0 = __shfl_sync(FULL, d, leader);
    const unsigned same = __ballot_sync(FULL, active && d == d0);
    if ((threadIdx.x & 31) == leader) atomicAdd(&hist[d0], (uint32_t)__popc(same));
    if (active && d != d0) atomicAdd(&hist[d], 1u);
}

// k-th smallest (0-based) of a multiset enumerated cooperatively by the CTA: `make()` returns a per-thread
// enumerator whose next(key) yields this thread's items (every thread is stepped `rounds` times so the
// warp votes in hist_add stay convergent).
template <class Make>
__device__ unsigned long long radix_select64(Make make, uint32_t rounds, uint32_t kth, uint32_t *hist,
                                             unsigned long long *sh_prefix, uint32_t *sh_rem) {
    const int tid = threadIdx.x;
    if (tid == 0) { *sh_prefix = 0ull; *sh_rem = kth; }
    __syncthreads();
    for (int level = 0; level < 8; level++) {
        const int shift = 56 - 8 * level;
        const unsigned long long himask = level == 0 ? 0ull : (~0ull << (shift + 8));
        const unsigned long long prefix = *sh_prefix;
        hist[tid] = 0;   // TS_THREADS == 256 bins
        __syncthreads();
        auto en = make();
        for (uint32_t it = 0; it < rounds; it++) {
            unsigned long long key = 0;
            bool active = en.next(key);
            active = active && (key & himask) == prefix;
            hist_add(hist, (uint32_t)(key >> shift) & 0xFFu, active);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rem = *sh_rem, d = 0;
            for (; d < 255; d++) {
                const uint32_t c = hist[d];
                if (rem < c) break;
                rem -= c;
            }
            *sh_prefix = prefix | ((unsigned long long)d << shift);
            *sh_rem = rem;
        }
        __syncthreads();
    }
    return *sh_prefix;
}

// All unordered pairs {i, j} of np points, circulant order: round d = 1..(np-1)/2 pairs i with (i+d) mod np; for even
// np a final half round d = np/2 with i < np/2.  Each pair is oriented (lower index first) exactly as the reference's
// nested loop (event_handling.cpp:67-75), so dx, dy and the signed zero of dy/dx are the reference's.
struct PairEnum {
    const double *x, *y;
    uint32_t np, d, i, full_rounds;
    __device__ bool next(unsigned long long &key) {
        while (i >= np) { i -= np; d++; }
        const bool in_full = d <= full_rounds;
        const bool in_half = (np % 2 == 0) && d == np / 2 && i < np / 2;
        bool ok = in_full || in_half;
        if (ok) {
            uint32_t j = i + d;
            if (j >= np) j -= np;
            const uint32_t lo = min(i, j), hi = max(i, j);
            key = order_key(dDiv(dSub(y[lo], y[hi]), dSub(x[lo], x[hi])));
        }
        i += TS_THREADS;
        return ok;
    }
};

struct IcptEnum {
    const double *x, *y;
    double slope;
    uint32_t np, i;
    __device__ bool next(unsigned long long &key) {
        const bool ok = i < np;
        if (ok) key = order_key(dSub(y[i], dMul(slope, x[i])));   // :83, not fused
        i += TS_THREADS;
        return ok;
    }
};

__global__ void __launch_bounds__(TS_THREADS) theil_sen_kernel(DnbBatchView v, DnbModelDev m, DnbTsArgs a) {
    __shared__ double sx[TS_MAXP], sy[TS_MAXP];
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long sh_prefix;
    __shared__ uint32_t sh_rem;
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
        sx[j] = dDiv(dSub(sig[i], shift), scale);          // :51
        sy[j] = m.mean[rk[i]];                              // :58
    }
    __syncthreads();

    // median slope: element ns/2 of the ascending sort of dy/dx over all i<j (:67-78)
    const uint32_t ns = np * (np - 1) / 2;
    const uint32_t full_rounds = (np - 1) / 2;
    const uint32_t items = full_rounds * np + ((np % 2 == 0) ? np : 0);   // enumeration span incl. the padded half round
    auto make_pairs = [&]() { return PairEnum{sx, sy, np, 1u, (uint32_t)tid, full_rounds}; };
    const double slope = key_to_double(
        radix_select64(make_pairs, (items + TS_THREADS - 1) / TS_THREADS, ns / 2, hist, &sh_prefix, &sh_rem));
    __syncthreads();

    // median intercept: element np/2 of y - slope*x (:81-87)
    auto make_icpt = [&]() { return IcptEnum{sx, sy, slope, np, (uint32_t)tid}; };
    const double icpt = key_to_double(
        radix_select64(make_icpt, (np + TS_THREADS - 1) / TS_THREADS, np / 2, hist, &sh_prefix, &sh_rem));

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
