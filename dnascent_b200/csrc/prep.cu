// prep.cu -- k-mer ranks, quantile (rough) scaling and event scaling.
//
// Replaces (reference, paths relative to /root/reference):
//   kmer2index + rank precompute      src/data_IO.cpp:129-141, src/event_handling.cpp:578-592
//   estimateScaling_quantiles         src/event_handling.cpp:510-541
//   quantileMedians                   src/event_handling.cpp:451-475   (full std::sort there; exact selection here)
//   linear_regression                 src/event_handling.cpp:478-507
//   x = (e.mean - shift)/scale        src/event_handling.cpp:130       (hoisted: depends on the event only)
#include "dnb_internal.cuh"

namespace {

// ---- ranks ------------------------------------------------------------------------------------------------------
// One CTA per (read, which) pair; thread i handles k-mers i, i+blockDim, ...  A k-mer is 9 bytes of a string the
// CTA walks contiguously, so the loads coalesce and hit L1 eight times out of nine.
__global__ void __launch_bounds__(256) ranks_kernel(DnbBatchView v, DnbModelDev m, double *mu_q, uint32_t *rank_ref) {
    const uint32_t r = v.order[blockIdx.x];
    const bool is_ref = blockIdx.y == 1;
    const uint64_t off = is_ref ? v.r_off[r] : v.q_off[r];
    const uint64_t len = (is_ref ? v.r_off[r + 1] : v.q_off[r + 1]) - off;
    if (len < DNB_K) return;
    const char *seq = (is_ref ? v.ref : v.query) + off;
    const uint64_t nk = len - DNB_K + 1;
    for (uint64_t i = threadIdx.x; i < nk; i += blockDim.x) {
        uint32_t rk = 0;
#pragma unroll
        for (int j = 0; j < DNB_K; j++) rk = rk * 4u + dnb_base_code(seq[i + j]);   // first base most significant
        if (is_ref) rank_ref[off + i] = rk;
        else mu_q[off + i] = m.mean[rk];
    }
}

// ---- exact multi-target order statistics (radix select, 4 levels of 8 bits over u32 keys) -------------------------
#define QS_THREADS 256
#define QS_NT 10   // nquantiles, event_handling.cpp:532

template <class KeyFn>
__device__ void select10(KeyFn key, uint32_t n, const uint32_t *target /*[10] smem*/, uint32_t *result /*[10] smem*/,
                         uint32_t *hist /*[10*256] smem*/, uint32_t *prefix /*[10] smem*/, uint32_t *remain /*[10] smem*/,
                         int *lead /*[10] smem*/) {
    const int tid = threadIdx.x;
    if (tid < QS_NT) { prefix[tid] = 0; remain[tid] = target[tid]; }
    __syncthreads();
    for (int level = 0; level < 4; level++) {
        const int shift = 24 - 8 * level;
        const uint32_t himask = level == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        // targets are ascending, so equal prefixes are adjacent: one histogram per distinct prefix
        if (tid < QS_NT) {
            int l = tid;
            while (l > 0 && prefix[l - 1] == prefix[tid]) l--;
            lead[tid] = l;
        }
        for (int i = tid; i < QS_NT * 256; i += QS_THREADS) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += QS_THREADS) {
            const uint32_t k = key(i);
            const uint32_t d = (k >> shift) & 0xFFu;
#pragma unroll
            for (int t = 0; t < QS_NT; t++)
                if (lead[t] == t && (k & himask) == prefix[t]) atomicAdd(&hist[t * 256 + d], 1u);
        }
        __syncthreads();
        if (tid < QS_NT) {
            const uint32_t *h = hist + lead[tid] * 256;
            uint32_t rem = remain[tid], d = 0;
            for (; d < 255; d++) {
                uint32_t c = h[d];
                if (rem < c) break;
                rem -= c;
            }
            prefix[tid] |= d << shift;
            remain[tid] = rem;
        }
        __syncthreads();
    }
    if (tid < QS_NT) result[tid] = prefix[tid];
    __syncthreads();
}

__global__ void __launch_bounds__(QS_THREADS) quantile_kernel(DnbBatchView v, DnbModelDev m, const uint32_t *rank_ref,
                                                              double *rough_shift, double *rough_scale) {
    __shared__ uint32_t hist[QS_NT * 256];
    __shared__ uint32_t target[QS_NT], result[QS_NT], prefix[QS_NT], remain[QS_NT];
    __shared__ int lead[QS_NT];
    __shared__ double sig_q[QS_NT], mod_q[QS_NT];
    const uint32_t r = v.order[blockIdx.x];
    const uint32_t E = v.n_events[r];
    const uint64_t roff = v.r_off[r];
    const uint64_t rlen = v.r_off[r + 1] - roff;
    if (E == 0 || rlen < DNB_K || v.status[r] != 0) return;   // undefined in the reference: status already says so
    const uint32_t Kref = (uint32_t)(rlen - DNB_K + 1);
    const int tid = threadIdx.x;

    // signal quantiles: keys = bit patterns of the (non-negative) float event means -> unsigned order == value order
    const float *em = v.ev_mean + v.ev_off[r];
    if (tid < QS_NT) {
        uint32_t n = E / QS_NT;                               // unsigned int n = size / nquantiles  (:467)
        target[tid] = ((uint32_t)tid * n + (uint32_t)(tid + 1) * n) / 2u;
    }
    __syncthreads();
    select10([&](uint32_t i) { return __float_as_uint(em[i]); }, E, target, result, hist, prefix, remain, lead);
    if (tid < QS_NT) sig_q[tid] = (double)__uint_as_float(result[tid]);
    __syncthreads();

    // model quantiles: keys = position of the k-mer's mean in the ascending sort of the whole table
    const uint32_t *rr = rank_ref + roff;
    if (tid < QS_NT) {
        uint32_t n = Kref / QS_NT;
        target[tid] = ((uint32_t)tid * n + (uint32_t)(tid + 1) * n) / 2u;
    }
    __syncthreads();
    select10([&](uint32_t i) { return m.mean_order[rr[i]]; }, Kref, target, result, hist, prefix, remain, lead);
    if (tid < QS_NT) mod_q[tid] = m.sorted_mean[result[tid]];
    __syncthreads();

    if (tid == 0) {
        // linear_regression(x = model quantiles, y = signal quantiles), sums in index order, nothing fused
        double sx = 0., sx2 = 0., sy = 0., sxy = 0.;
        for (int i = 0; i < QS_NT; i++) {
            sx = dAdd(sx, mod_q[i]);
            sx2 = dAdd(sx2, dMul(mod_q[i], mod_q[i]));
            sy = dAdd(sy, sig_q[i]);
            sxy = dAdd(sxy, dMul(mod_q[i], sig_q[i]));
        }
        const double n = (double)QS_NT;
        double slope = dDiv(dSub(dMul(n, sxy), dMul(sx, sy)), dSub(dMul(n, sx2), dMul(sx, sx)));
        double icpt = dDiv(dSub(sy, dMul(slope, sx)), n);
        rough_shift[r] = icpt;   // s.shift = scalings.second (:537)
        rough_scale[r] = slope;  // s.scale = scalings.first  (:538)
    }
}

// x_e = (mean - shift) / scale, event_handling.cpp:130
__global__ void __launch_bounds__(256) scale_events_kernel(DnbBatchView v, const double *rough_shift,
                                                           const double *rough_scale, double *x_e) {
    const uint32_t r = v.order[blockIdx.x];
    if (v.status[r] != 0) return;
    const uint32_t E = v.n_events[r];
    const double sh = rough_shift[r], sc = rough_scale[r];
    const float *em = v.ev_mean + v.ev_off[r];
    double *x = x_e + v.ev_off[r];
    for (uint32_t i = threadIdx.x; i < E; i += blockDim.x) x[i] = dDiv(dSub((double)em[i], sh), sc);
}

}  // namespace

void dnb_launch_ranks(const DnbBatchView &v, const DnbModelDev &m, double *mu_q, uint32_t *rank_ref, cudaStream_t s) {
    if (v.n_reads == 0) return;
    ranks_kernel<<<dim3(v.n_reads, 2), 256, 0, s>>>(v, m, mu_q, rank_ref);
}

void dnb_launch_quantile_scaling(const DnbBatchView &v, const DnbModelDev &m, const uint32_t *rank_ref,
                                 double *rough_shift, double *rough_scale, cudaStream_t s) {
    if (v.n_reads == 0) return;
    quantile_kernel<<<v.n_reads, QS_THREADS, 0, s>>>(v, m, rank_ref, rough_shift, rough_scale);
}

void dnb_launch_scale_events(const DnbBatchView &v, const double *rough_shift, const double *rough_scale, double *x_e,
                             cudaStream_t s) {
    if (v.n_reads == 0) return;
    scale_events_kernel<<<v.n_reads, 256, 0, s>>>(v, rough_shift, rough_scale, x_e);
}
