// features.cu -- the DNN input tensors built on the device from eventalign's records.
//
// Replaces (reference, paths relative to /root/reference):
//   read::addSignal                       src/reads.h:288-300    called per raw sample from eventalign, alignment.cpp:723
//   AlignedPosition::makeSignalFeature    src/reads.h:147-172    first RAWDEPTH = 20 scaled samples, zero padded
//   AlignedPosition::getCoreIndex / getResidualIndex   src/reads.h:109-138
//   read::makeSignalTensor / makeCoreSequenceTensor / makeResidualSequenceTensor / getReferenceCoords /
//   getReferenceIndices / getQueryIndices / getAlignmentQuality          src/reads.h:305-452
// Consumer: runCNN (src/detect.cpp:586-649) feeds exactly these vectors to TensorFlow -- that stays the reference's.
//
// What the reference does, restated as data flow.  eventalign walks the read's windows; every match-state event adds
// its raw samples, scaled with the read's final scalings, to the map entry of its reference coordinate (unless that
// coordinate already has a call, alignment.cpp:711).  The next window starts one past the last match of the current
// one (alignment.cpp:741), so over the whole read the match records come with NON-DECREASING reference positions and
// the records of one position are consecutive among the match records.  The std::map therefore is "runs of equal
// ref_pos among the M records", and both iteration orders the tensor builders use (begin->end on "fwd", rbegin->rend
// on "rev", where the coordinate decreases with the position) are ascending ref_pos: output row o is the o-th run.
//
// Kernel: one CTA per read (persistent CTAs, atomic read counter), 256 records per chunk.
//   1. exclusive max-scan of (ref_pos + 1) over the M records -> is this record the head of a run?  (a decreasing
//      position, which eventalign cannot produce, marks the read DNB_READ_UNDEFINED instead of emitting garbage)
//   2. every head walks its run and stages WHICH raw sample fills each of its 20 columns (integer work) in a
//      shared-memory row (stride 21 words: conflict free)
//   3. exclusive sum-scan of the kept heads -> output row; then the whole CTA converts: element j of the chunk's
//      contiguous [rows][20] block is (float)((raw - shift) / scale) with IEEE double subtract / divide
//      (alignment.cpp:709, reads.h:154) or zero padding, raw recomputed from the int16 DAC with pod5.cpp:60's float
//      expression where that was shipped -- every thread does the same number of divisions and the 80-byte rows
//      leave as one fully coalesced store per chunk; the six per-position scalars are written by the head threads.
// HBM traffic per position: 16 B of record + ~20 samples read, 108 B written -- a streaming kernel.
#include <cub/block/block_scan.cuh>
#include "dnb_internal.cuh"
#include "../../include/dnascent_b200.h"

#define FT_THREADS 256
#define FT_ROW (DNB_RAWDEPTH + 1)
#define FT_NONE 0xffffffffu

namespace {

struct MaxOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// r.refCoordToCalls.count(coord) on the sorted key list
__device__ __forceinline__ bool called_contains(const uint32_t *v, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t m = (lo + hi) >> 1;
        if (v[m] < x) lo = m + 1; else hi = m;
    }
    return lo < n && v[lo] == x;
}

__global__ void __launch_bounds__(FT_THREADS) features_kernel(DnbFeatArgs a) {
    typedef cub::BlockScan<uint32_t, FT_THREADS> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t rows[FT_THREADS * FT_ROW];
    __shared__ uint16_t src_of[FT_THREADS];
    __shared__ uint32_t s_read;
    __shared__ int s_bad;
    const unsigned tid = threadIdx.x;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_read = atomicAdd(a.next_read, 1u); s_bad = 0; }
        __syncthreads();
        if (s_read >= a.n_reads) break;
        const uint32_t r = a.order ? a.order[s_read] : s_read;
        if (a.status[r] != DNB_READ_OK) { if (tid == 0) a.n_pos[r] = 0; continue; }

        const dnb_eventalign_rec *recs = a.recs + a.rec_off[r];
        const uint32_t nrec = a.n_rec[r];
        const char *ref = a.ref + a.ref_off[r];
        const int32_t *r2q = a.r2q + a.ref_off[r];
        const uint32_t *es = a.ev_start + a.ev_off[r] + r;             // n_events + 1 entries
        const uint64_t raw0 = a.raw_off[r];
        const bool dac = a.raw_kind[r] != 0;
        const float dac_off = dac ? a.dac_offset[r] : 0.f, dac_scl = dac ? a.dac_scale[r] : 1.f;
        const double shift = a.shift[r], scale = a.scale[r];
        const bool rev = a.is_reverse[r] != 0;
        const uint32_t ref_start = a.ref_start[r], ref_end = a.ref_end[r];
        const uint32_t *called = a.called + a.called_off[r];
        const uint32_t n_called = (uint32_t)(a.called_off[r + 1] - a.called_off[r]);
        const uint64_t out0 = a.pos_off[r];
        const uint32_t cap = (uint32_t)(a.pos_off[r + 1] - a.pos_off[r]);

        uint32_t out_base = 0;       // rows emitted by earlier chunks
        uint32_t carry_max = 0;      // (ref_pos + 1) of the last M record of earlier chunks, 0 = none
        bool overflow = false;

        for (uint32_t chunk = 0; chunk < nrec; chunk += FT_THREADS) {
            const uint32_t q = chunk + tid;
            dnb_eventalign_rec rec = {0u, 0u, 0, 0u};
            if (q < nrec) rec = recs[q];
            const bool is_m = q < nrec && rec.label == DNB_EA_MATCH;
            const uint32_t key = is_m ? rec.ref_pos + 1u : 0u;
            uint32_t prev, chunk_max;
            Scan(scan_tmp).ExclusiveScan(key, prev, 0u, MaxOp(), chunk_max);
            prev = prev > carry_max ? prev : carry_max;
            if (is_m && key < prev) s_bad = 1;                               // decreasing position: not eventalign output
            const bool head = is_m && key > prev;

            // (2) the head's run -> its shared-memory row
            uint32_t nsig = 0;
            bool keep = false;
            uint32_t coord = 0;
            if (head) {
                coord = rev ? ref_end - rec.ref_pos - DNB_K / 2 - 1u : ref_start + rec.ref_pos + DNB_K / 2;   // alignment.cpp:646-648, 690-697
                keep = !called_contains(called, n_called, coord);                                               // :711
            }
            if (keep) {
                // integer work only: which raw sample fills each of the 20 columns (FT_NONE = zero padding); the
                // floating-point conversion is done by the whole CTA below
                uint32_t *row = rows + tid * FT_ROW;
                for (uint32_t j = q; j < nrec && nsig < DNB_RAWDEPTH; j++) {
                    dnb_eventalign_rec rj = rec;
                    if (j != q) {
                        rj = recs[j];
                        if (rj.label != DNB_EA_MATCH) continue;
                        if (rj.ref_pos != rec.ref_pos) break;
                    }
                    const uint32_t s0 = es[rj.event], s1 = es[rj.event + 1];
                    for (uint32_t t = s0; t < s1 && nsig < DNB_RAWDEPTH; t++) row[nsig++] = t;
                }
                for (uint32_t t = nsig; t < DNB_RAWDEPTH; t++) row[t] = FT_NONE;                 // reads.h:162-168
                keep = nsig > 0;                                                                    // addSignal is per sample
            }

            // (3) output rows
            uint32_t local, n_local;
            __syncthreads();                                                   // scan_tmp reuse
            Scan(scan_tmp).ExclusiveSum(keep ? 1u : 0u, local, n_local);
            const uint32_t o = out_base + local;
            if (keep) {
                src_of[local] = (uint16_t)tid;
                if (o < cap) {
                    const char *km = ref + rec.ref_pos;
                    uint32_t c = 0;
#pragma unroll
                    for (int i = 2; i < 7; i++) c = c * 4u + dnb_base_code(km[i]);                        // reads.h:109-121
                    const uint32_t rs = ((dnb_base_code(km[0]) * 4u + dnb_base_code(km[1])) * 4u + dnb_base_code(km[7])) * 4u +
                                        dnb_base_code(km[8]);                                              // reads.h:122-135
                    const uint32_t iref = rec.ref_pos + DNB_K / 2;                                         // alignment.cpp:701
                    a.core[out0 + o] = (float)(c + 1u);
                    a.residual[out0 + o] = (float)(rs + 1u);
                    a.coords[out0 + o] = coord;
                    a.ref_index[out0 + o] = iref;
                    a.query_index[out0 + o] = (uint32_t)r2q[iref];
                    a.quality[out0 + o] = rec.indel_score;
                }
            }
            __syncthreads();
            uint32_t n_write = n_local;
            if (out_base + n_local > cap) { overflow = true; n_write = cap > out_base ? cap - out_base : 0u; }
            float *dst = a.signal + (out0 + out_base) * DNB_RAWDEPTH;
            for (uint32_t j = tid; j < n_write * DNB_RAWDEPTH; j += FT_THREADS) {
                const uint32_t row = j / DNB_RAWDEPTH, col = j - row * DNB_RAWDEPTH;
                const uint32_t t = rows[src_of[row] * FT_ROW + col];
                float v = 0.f;
                if (t != FT_NONE) {
                    float pa;
                    if (dac) pa = fMul(fAdd((float)a.raw_i16[raw0 + t], dac_off), dac_scl);       // pod5.cpp:60
                    else pa = a.raw_f32[raw0 + t];
                    v = d2f(dDiv(dSub((double)pa, shift), scale));                                // alignment.cpp:709, reads.h:154
                }
                dst[j] = v;
            }
            out_base += n_local;
            carry_max = chunk_max > carry_max ? chunk_max : carry_max;
            __syncthreads();
        }
        if (tid == 0) {
            if (s_bad) { a.status[r] = DNB_READ_UNDEFINED; a.n_pos[r] = 0; }
            else if (overflow) { a.status[r] = DNB_READ_OVERFLOW; a.n_pos[r] = 0; }
            else a.n_pos[r] = out_base;
        }
    }
}

}  // namespace

unsigned dnb_features_grid(int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, features_kernel, FT_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    return (unsigned)(sms * per_sm);
}

void dnb_launch_features(const DnbFeatArgs &a, unsigned grid, cudaStream_t s) {
    if (a.n_reads == 0) return;
    if (grid > a.n_reads) grid = a.n_reads;
    features_kernel<<<grid, FT_THREADS, 0, s>>>(a);
}
