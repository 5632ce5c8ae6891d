// pack.cu -- the compact wire formats either side of the pipeline (what crosses PCIe), and their device-side
// expansion / production.  Nothing here changes a value: every kernel is a re-encoding that the host (or the
// next kernel) inverts exactly.
//
//   host -> device   queryToRef as runs              dnb_q2r_run        (reference: std::map from parseCigar,
//                                                                        src/htsInterface.cpp:59-157; dense int32 costs
//                                                                        4 B per query base, a `{L}M` read is ONE run)
//   device -> host   event table, delta coded        u8 length per event + f32 mean (+ in-order escapes >= 255)
//                    alignment as a monotone path    2 bits per step: consecutive eventAlignment pairs differ by
//                                                    (event,kmer) in {(+1,+1), (+1,0), (0,+1)} -- the three trace moves
//                                                    of the backtrace (src/event_handling.cpp:160-162, 347-413)
//
// Layouts are documented next to dnb_read_result in include/dnascent_b200.h.
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

namespace {

// ---- queryToRef runs -> dense int32 (4 B per query base, in HBM only) --------------------------------------------
// One warp per run; the dense array was set to -1 before.
__global__ void __launch_bounds__(128) expand_q2r_kernel(const dnb_q2r_run *runs, const uint64_t *run_q_base,
                                                         const uint32_t *run_q_len, uint64_t n_runs, int32_t *q2r) {
    const uint64_t g = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n_runs) return;
    const dnb_q2r_run run = runs[g];
    const uint32_t qlen = run_q_len[g];
    if (run.q_start >= qlen) return;
    const uint32_t len = min(run.len, qlen - run.q_start);      // a run never writes outside its read
    int32_t *dst = q2r + run_q_base[g] + run.q_start;
    for (uint32_t i = lane; i < len; i += 32) dst[i] = run.r_start + (run.stride ? (int32_t)i : 0);
}

// ---- events: how many lengths need an escape (>= 255 samples), per read ------------------------------------------
__global__ void __launch_bounds__(256) count_long_events_kernel(DnbBatchView v, uint32_t *n_long) {
    __shared__ uint32_t acc;
    const uint32_t r = blockIdx.x;
    if (threadIdx.x == 0) acc = 0;
    __syncthreads();
    const uint32_t n = (v.status[r] == DNB_READ_OVERFLOW) ? 0u : v.n_events[r];
    const uint32_t *ss = v.ev_start + v.ev_off[r] + r;
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) c += (ss[i + 1] - ss[i]) >= 255u;
    if (c) atomicAdd(&acc, c);
    __syncthreads();
    if (threadIdx.x == 0) n_long[r] = acc;
}

// ---- events: capacity-strided (u32 start, f32 mean) slots -> dense (u8 length, f32 mean) + escapes in event order --
#define CE_THREADS 256
__global__ void __launch_bounds__(CE_THREADS) compact_events8_kernel(DnbBatchView v, const uint64_t *dense_off,
                                                                     const uint64_t *esc_off, uint8_t *out_len8,
                                                                     float *out_mean, uint32_t *out_esc,
                                                                     uint32_t *out_first) {
    __shared__ uint32_t warp_cnt[CE_THREADS / 32];
    __shared__ uint32_t base;
    const uint32_t r = blockIdx.x;
    const uint32_t n = (uint32_t)(dense_off[r + 1] - dense_off[r]);
    const uint32_t *ss = v.ev_start + v.ev_off[r] + r;
    if (threadIdx.x == 0) { base = 0; out_first[r] = n ? ss[0] : 0u; }
    if (n == 0) return;
    const float *ms = v.ev_mean + v.ev_off[r];
    uint8_t *dl = out_len8 + dense_off[r];
    float *dm = out_mean + dense_off[r];
    uint32_t *de = out_esc + esc_off[r];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += CE_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t d = 0;
        if (i < n) { d = ss[i + 1] - ss[i]; dm[i] = ms[i]; }
        const bool esc = i < n && d >= 255u;
        if (i < n) dl[i] = esc ? (uint8_t)255 : (uint8_t)d;
        const unsigned bal = __ballot_sync(0xffffffffu, esc);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = base;
        for (int w = 0; w < wid; w++) before += warp_cnt[w];
        if (esc) de[before + __popc(bal & ((1u << lane) - 1u))] = d;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < CE_THREADS / 32; w++) t += warp_cnt[w];
            base += t;
        }
        __syncthreads();
    }
}

// ---- alignment: reversed (event, kmer) pairs -> first pair + 2-bit forward steps -----------------------------------
// forward pair t = rev[n-1-t]; step t (t = 0 .. n-2) leads from pair t to pair t+1 and is coded like the trace move
// that produced it: DNB_FROM_D (+1,+1), DNB_FROM_U (+1,0), DNB_FROM_L (0,+1).  Byte q of a read holds steps 4q..4q+3,
// step t at bits 2*(t&3).  A pair of steps that is none of the three (cannot happen for a backtrace) sets bad[r].
__global__ void __launch_bounds__(128) compact_steps_kernel(const uint64_t *al_off, const uint32_t *al_pairs_rev,
                                                            const uint32_t *n_align, const uint64_t *step_off,
                                                            uint8_t *out_steps, uint32_t *out_first, uint32_t *bad) {
    const uint32_t r = blockIdx.x;
    const uint32_t n = n_align[r];
    const uint2 *src = reinterpret_cast<const uint2 *>(al_pairs_rev) + al_off[r];
    if (threadIdx.x == 0) {
        const uint2 f = n ? src[n - 1] : make_uint2(0u, 0u);
        out_first[2 * r] = f.x; out_first[2 * r + 1] = f.y;
    }
    if (n < 2) return;
    uint8_t *dst = out_steps + step_off[r];
    const uint32_t n_steps = n - 1, n_bytes = (n_steps + 3) / 4;
    bool wrong = false;
    for (uint32_t q = threadIdx.x; q < n_bytes; q += blockDim.x) {
        uint32_t byte = 0;
        uint2 a = src[n - 1 - 4 * q];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t t = 4 * q + j;
            if (t >= n_steps) break;
            const uint2 b = src[n - 2 - t];
            const uint32_t de = b.x - a.x, dk = b.y - a.y;
            uint32_t code;
            if (de == 1u && dk == 1u) code = DNB_FROM_D;
            else if (de == 1u && dk == 0u) code = DNB_FROM_U;
            else if (de == 0u && dk == 1u) code = DNB_FROM_L;
            else { code = 3u; wrong = true; }
            byte |= code << (2 * j);
            a = b;
        }
        dst[q] = (uint8_t)byte;
    }
    if (wrong) bad[r] = 1u;
}

// zero the padding between a read's last sample and the next read's (32-element aligned) start
__global__ void __launch_bounds__(128) zero_padding_kernel(DnbBatchView v, uint32_t esz) {
    const uint32_t r = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= v.n_reads) return;
    const uint64_t lo = (v.raw_off[r] + v.n_samples[r]) * esz, hi = v.raw_off[r + 1] * esz;
    uint8_t *p = v.raw_i16 ? (uint8_t *)v.raw_i16 : (uint8_t *)v.raw_f32;
    for (uint64_t i = lo + (threadIdx.x & 31); i < hi; i += 32) p[i] = 0;
}

}  // namespace

void dnb_launch_expand_q2r(const dnb_q2r_run *runs, const uint64_t *run_q_base, const uint32_t *run_q_len, uint64_t n_runs,
                           int32_t *q2r, cudaStream_t s) {
    if (n_runs == 0) return;
    expand_q2r_kernel<<<(unsigned)((n_runs + 3) / 4), 128, 0, s>>>(runs, run_q_base, run_q_len, n_runs, q2r);
}

void dnb_launch_count_long_events(const DnbBatchView &v, uint32_t *n_long, cudaStream_t s) {
    if (v.n_reads == 0) return;
    count_long_events_kernel<<<v.n_reads, 256, 0, s>>>(v, n_long);
}

void dnb_launch_compact_events8(const DnbBatchView &v, const uint64_t *dense_off, const uint64_t *esc_off, uint8_t *out_len8,
                                float *out_mean, uint32_t *out_esc, uint32_t *out_first, cudaStream_t s) {
    if (v.n_reads == 0) return;
    compact_events8_kernel<<<v.n_reads, CE_THREADS, 0, s>>>(v, dense_off, esc_off, out_len8, out_mean, out_esc, out_first);
}

void dnb_launch_compact_steps(const DnbBatchView &v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                              const uint32_t *n_align, const uint64_t *step_off, uint8_t *out_steps, uint32_t *out_first,
                              uint32_t *bad, cudaStream_t s) {
    if (v.n_reads == 0) return;
    compact_steps_kernel<<<v.n_reads, 128, 0, s>>>(al_off, al_pairs_rev, n_align, step_off, out_steps, out_first, bad);
}

void dnb_launch_zero_padding(const DnbBatchView &v, uint32_t esz, cudaStream_t s) {
    if (v.n_reads == 0) return;
    zero_padding_kernel<<<(v.n_reads + 3) / 4, 128, 0, s>>>(v, esz);
}
