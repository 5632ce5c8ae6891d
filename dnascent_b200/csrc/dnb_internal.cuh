// dnb_internal.cuh -- shared device/host declarations of libdnascent_b200 (sm_100a only).
//
// Arithmetic discipline: the reference is x86-64 baseline code (no FMA, float ops in
// SSE single, double ops in SSE double), so every operation whose rounding is observable is written with an
// explicit IEEE round-to-nearest intrinsic (__dadd_rn, __fmul_rn, ...) -- these are never contracted into
// FMAs by nvcc -- and the library is additionally built with -fmad=false -prec-div=true -prec-sqrt=true
// -ftz=false.  Where an FMA is used on purpose its product is exact and the comment says why.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

#define DNB_BW 100          // AdaptiveBanded_Params.bandwidth, src/config.h:41
#define DNB_K 9
#define DNB_TRACE_ROW 32    // bytes per band in HBM: 128 event slots x 2-bit trace code
#define DNB_CELLS_PER_LANE 4

enum { DNB_FROM_D = 0, DNB_FROM_U = 1, DNB_FROM_L = 2 };   // event_handling.cpp:160-162

// ---- typed arithmetic helpers: one rounding each, exactly as written -------------------------------------------
__device__ __forceinline__ double dAdd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dSub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dMul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dDiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float fAdd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fSub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fMul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fDiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float d2f(double a) { return __double2float_rn(a); }

// base code of the reference alphabet A=0 T=1 G=2 C=3, anything else 0 (src/data_IO.cpp:131-137, quirk Q11)
__host__ __device__ __forceinline__ uint32_t dnb_base_code(char c) {
    return c == 'T' ? 1u : c == 'G' ? 2u : c == 'C' ? 3u : 0u;
}

// ---- per-batch device view handed to the kernels ---------------------------------------------------------------
struct DnbDetector {
    uint32_t w1, w2;
    float thr1, thr2, peak_height;
};

struct DnbBatchView {
    uint32_t n_reads;
    const uint32_t *order;        // processing order (longest first); kernels map slot -> read = order[slot]
    // inputs (SoA, concatenated; offsets in elements)
    const float *raw_f32;         // or nullptr
    const int16_t *raw_i16;       // or nullptr
    const float *dac_offset, *dac_scale;   // per read (i16 input only)
    const uint64_t *raw_off;      // [R+1], each read starts 32-element aligned
    const uint32_t *n_samples;    // [R]
    const char *query;            // concatenated
    const char *ref;
    const uint64_t *q_off;        // [R+1]
    const uint64_t *r_off;        // [R+1]
    const int32_t *q2r;           // indexed like query
    // segmentation outputs
    uint32_t *et_n;               // [R] scrappie event count (0 = undefined)
    uint32_t *n_events;           // [R] r.events.size()
    const uint64_t *ev_off;       // [R+1] event slot offsets (capacity per read)
    uint32_t *ev_start;           // [ev_off[R] + R]  (n_events+1 entries per read, slot base ev_off[r] + r)
    float *ev_mean;               // [ev_off[R]]
    int *status;                  // [R] DNB_READ_*
    // optional full scrappie table (dnb_detect_events)
    uint64_t *et_start;
    float *et_length, *et_mean, *et_stdv;
};

// ---- kernel launchers (host side, defined next to each kernel) -----------------------------------------------------
struct DnbModelDev {
    const double *mean;           // [4^9]
    const double *stdv;           // [4^9] or nullptr (static 0.14)
    const uint32_t *mean_order;   // [4^9] position of mean[rank] in the ascending sort of all means
    const double *sorted_mean;    // [4^9]
};

// ---- tiled segmentation workspace (seg.cu) ----------------------------------------------------------------------
#define DNB_SEG_TILE 512u       // samples per lane in the tile kernel
#define DNB_SEG_HALO 64         // samples the detectors run before their tile, from a fresh state
#define DNB_SEG_CK 64u          // checkpoint spacing of the serial (sum, sumsq) chains
#define DNB_SEG_PEAK_CAP 224u   // peak slots per tile (observed ~0.2 peaks/sample -> ~100 per tile)

// detector state at a tile boundary, normalised so that equal records <=> identical future behaviour
struct __align__(8) SegBoundary {
    int s_pos, l_pos;
    float s_val, l_val;
    uint32_t masked;   // long detector's masked_to if it still masks a position >= the boundary, else 0
    uint32_t valid;    // bit0 short.valid, bit1 long.valid
};
struct DnbSegTiles {
    uint32_t n_tiles;
    const uint32_t *tile_off;     // [R+1] first global tile of each read
    const uint32_t *tile_read;    // [n_tiles]
    const uint64_t *ck_off;       // [R+1] checkpoint slot offsets
    double *ck_sum, *ck_sq;       // [ck_off[R]]
    double *tot_sum;              // [R] sums[N]
    uint32_t *pk_pos;             // [n_tiles * DNB_SEG_PEAK_CAP]
    double *pk_sum;               // [n_tiles * DNB_SEG_PEAK_CAP] sums[peak]
    uint32_t *pk_count;           // [n_tiles]
    SegBoundary *b_start, *b_end; // [n_tiles] detector state assumed at the tile start / reached at its end
    uint32_t *tile_prefix;        // [n_tiles] peaks of the read before this tile
    uint32_t *tile_prev_pos;      // [n_tiles] last peak before this tile (0 if none) ...
    double *tile_prev_sum;        // ... and sums[] there
    uint32_t *redo;               // [R] 1 = redo this read with the serial kernel
};
#define DNB_SEG_BOUNDARY_BYTES sizeof(SegBoundary)

void dnb_launch_segmentation_serial(const DnbBatchView &v, DnbDetector det, const uint32_t *only_flagged, cudaStream_t s);
// after_checkpoint / after_tiles: optional events recorded behind the checkpoint (or scan) and the tile kernel
void dnb_launch_segmentation_tiled(const DnbBatchView &v, DnbDetector det, const DnbSegTiles &t, cudaStream_t s,
                                   cudaEvent_t after_checkpoint = nullptr, cudaEvent_t after_tiles = nullptr);
// the exact (sum, sumsq) checkpoints as a streaming warp scan (seg_scan.cu); DNB_SEG_PARITY_SCAN=0 selects the
// one-lane-per-read checkpoint kernel of seg.cu instead
bool dnb_seg_parity_scan_enabled(void);
cudaError_t dnb_launch_seg_parity_scan(const DnbBatchView &v, const DnbSegTiles &t, cudaStream_t s);
void dnb_launch_ranks(const DnbBatchView &v, const DnbModelDev &m, double *mu_q, uint32_t *rank_ref, cudaStream_t s);
void dnb_launch_quantile_scaling(const DnbBatchView &v, const DnbModelDev &m, const uint32_t *rank_ref,
                                 double *rough_shift, double *rough_scale, cudaStream_t s);
void dnb_launch_scale_events(const DnbBatchView &v, const double *rough_shift, const double *rough_scale, double *x_e,
                             cudaStream_t s);

struct DnbDpArgs {
    const double *x_e;            // indexed like ev_mean
    const double *mu_q;           // indexed like query (K entries used per read)
    const double *lp;             // [R][4] lp_skip, lp_stay, lp_step, lp_trim (host glibc, event_handling.cpp:174-183)
    double emit_const;            // (double)(float)log(1/sqrt(2pi)) - log(0.14)
    double inv_sigma;             // correctly rounded 1/0.14 for the guarded reciprocal multiply
    const uint64_t *band_off;     // [R+1] band-row offsets
    uint8_t *trace;               // [band_off[R]] rows of DNB_TRACE_ROW bytes: 2-bit code of slot s at bits 2*(s&3) of byte s>>2
    uint32_t *moves;              // [band_off[R]/32 + R + 1] move bits (1 = right), read r starts at band_off[r]/32 + r
    uint32_t *rcum;               // same indexing: right moves in the bands before each 32-band word
    int32_t *end_event;           // [R] event index of the best end cell, -1 if none
    int32_t *end_ll_event;        // [R] band_lower_left.event_idx of that band
    float *end_score;             // [R]
    unsigned long long *cells;    // [1] total DP cells filled (the reference's `fills`)
};

struct DnbBtArgs {
    DnbDpArgs dp;
    const uint32_t *rank_ref;     // indexed like ref
    const uint64_t *al_off;       // [R+1] alignment slot offsets (capacity n_bands per read)
    uint32_t *al_pairs_rev;       // [2*al_off[R]] (event,kmer) in backtrace order
    uint32_t *n_align;            // [R]
    const uint64_t *cl_off;       // [R+1] cleaned slot offsets (capacity K per read)
    double *cl_signal;
    uint32_t *cl_rank;
    uint32_t *n_cleaned;          // [R]
    double *avg_log_emission;     // [R]
    int *spanned, *max_gap;       // [R]
    double min_avg_log_emission;
    int max_gap_threshold;
    unsigned long long *phase_cycles;   // [2] summed warp cycles spent in the band fill / in the backtrace
};
// mode 0: band fill + backtrace in one launch; 1: band fill only; 2: backtrace only
void dnb_launch_align(const DnbBatchView &v, const DnbBtArgs &a, int mode, cudaStream_t s);

struct DnbTsArgs {
    const uint64_t *cl_off;
    const double *cl_signal;
    const uint32_t *cl_rank;
    const uint32_t *n_cleaned;
    const double *rough_shift, *rough_scale;
    double *shift, *scale;        // [R] refined
    int mode;                     // 0: approximate-quotient selection certified by exact divisions (default)
                                  // 1: the general exact digit search only (cross-check, DNB_TS_MODE=1)
    // reads with a 0/0 slope (theilsen.cu, theil_sen_nan_follow_kernel); nan_list == nullptr: no NaN path (the "NaN last" answer)
    uint32_t *nan_count;          // [1]
    uint32_t *nan_list;           // [8 * nan_cap]: 4 words per entry (read, status before Theil-Sen, action, -), then 2 * nan_cap doubles
                                  // (the slope read off the sorted range; the rank ns/2 - 1 slope)
    uint32_t nan_cap;
    double *nan_scratch;          // [nan_slots * DNB_TS_NAN_SLOT_DOUBLES]
    uint32_t nan_slots;           // CTAs of the NaN kernel, one scratch slot each
};
#define DNB_TS_MAX_SLOPES 499500ul                              /* 1000 * 999 / 2 */
#define DNB_TS_NAN_SLOT_DOUBLES (2ul * DNB_TS_MAX_SLOPES)        /* the slopes + two int position lists of the same length */
#define DNB_TS_NAN_SLOTS 16u
#define DNB_TS_NAN_CAP 1024u
void dnb_launch_theil_sen(const DnbBatchView &v, const DnbModelDev &m, const DnbTsArgs &a, cudaStream_t s);

void dnb_launch_compact_alignment(const DnbBatchView &v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                                  const uint32_t *n_align, const uint64_t *out_off, uint32_t *out_pairs,
                                  cudaStream_t s);

// capacity-strided event slots -> dense (start[n+1], mean[n]) per read at dense_off[r] (+ r for the starts)
void dnb_launch_compact_events(const DnbBatchView &v, const uint64_t *dense_off, uint32_t *out_start, float *out_mean,
                               cudaStream_t s);

// ---- compact wire formats (pack.cu) --------------------------------------------------------------------------------
struct dnb_q2r_run;
// runs -> dense q2r (already set to -1); run g belongs to the read whose query starts at run_q_base[g] and has run_q_len[g] bases
void dnb_launch_expand_q2r(const dnb_q2r_run *runs, const uint64_t *run_q_base, const uint32_t *run_q_len, uint64_t n_runs,
                           int32_t *q2r, cudaStream_t s);
// per read: number of events of >= 255 samples (the escapes of the u8 length coding)
void dnb_launch_count_long_events(const DnbBatchView &v, uint32_t *n_long, cudaStream_t s);
// capacity-strided event slots -> dense (u8 length, f32 mean) at dense_off[r], escapes at esc_off[r], event_start[0] in out_first[r]
void dnb_launch_compact_events8(const DnbBatchView &v, const uint64_t *dense_off, const uint64_t *esc_off, uint8_t *out_len8,
                                float *out_mean, uint32_t *out_esc, uint32_t *out_first, cudaStream_t s);
// backtrace-order pairs -> first pair (out_first[2r], [2r+1]) + 2-bit forward steps at byte offset step_off[r]
void dnb_launch_compact_steps(const DnbBatchView &v, const uint64_t *al_off, const uint32_t *al_pairs_rev,
                              const uint32_t *n_align, const uint64_t *step_off, uint8_t *out_steps, uint32_t *out_first,
                              uint32_t *bad, cudaStream_t s);
// zero the alignment padding after every read's last sample (direct per-read DMA leaves it unwritten)
void dnb_launch_zero_padding(const DnbBatchView &v, uint32_t esz, cudaStream_t s);

void dnb_launch_hmm_forward(const double *obs, const uint64_t *obs_off, const char *seq, const double *shift,
                            const double *scale, const double *epb, size_t n_sites, uint32_t window,
                            const DnbModelDev &unl, const DnbModelDev &ana, double *out_analogue, double *out_thymidine,
                            cudaStream_t s);

// ---- resident analogue stage (hmm.cu): llAcrossRead's site gathering + both forward passes per site -----------------
struct DnbLlrArgs {
    uint32_t window;              // windowLength of llAcrossRead (12)
    const int32_t *r2q;           // dense refToQuery, indexed like ref (an absent key reads as 0)
    const uint8_t *is_reverse;    // [R]
    const uint64_t *al_off;       // [R+1] alignment slot offsets
    const uint32_t *al_pairs_rev; // (event, k-mer) pairs in backtrace order
    const uint32_t *n_align;      // [R]
    const double *shift, *scale;  // [R] r.scalings
    // per reference base (capacity), filled by the site kernel: the read's T positions in ascending order and, per
    // site, the alignment index range [j_begin, j_end) of its events (j_end < 0: -(j_end + 1), visited downwards)
    uint32_t *poi;
    int32_t *j_begin, *j_end;
    uint32_t *n_poi;              // [R]
    const uint64_t *poi_off;      // [R+1] exclusive prefix of n_poi (host), visit order = ascending site index
    unsigned long long *next_site;   // work counter (zeroed before the forward launch)
    // per site (dense, poi_off indexing)
    uint32_t *out_pos, *out_n_events;
    double *out_a, *out_t;
};
void dnb_launch_llr_sites(const DnbBatchView &v, const DnbLlrArgs &a, cudaStream_t s);
void dnb_launch_llr_forward(const DnbBatchView &v, const DnbLlrArgs &a, uint64_t n_sites_total, const DnbModelDev &unl,
                            const DnbModelDev &ana, int device, cudaStream_t s);

// ---- eventalign (eventalign.cu): windowed Viterbi re-alignment, SURVEY s.8 row f1 -----------------------------------
struct dnb_eventalign_rec;
struct DnbEaArgs {
    uint32_t n_reads;
    uint32_t window;              // totalWindowLength (Global_Config.windowLength_align = 50)
    uint32_t t_max;               // observations per window the scratch rows hold
    const uint64_t *ref_off;      // [R+1] offsets into ref / r2q
    const char *ref;              // concatenated referenceSeqMappedTo
    const int32_t *r2q;           // dense refToQuery, indexed like ref
    const uint64_t *al_off;       // [R+1] offsets into pairs
    const uint2 *pairs;           // (event, k-mer) = r.eventAlignment
    const uint64_t *ev_off;       // [R+1] offsets into ev_mean
    const float *ev_mean;         // r.events[j].mean
    const double *shift, *scale;  // [R] r.scalings
    const double *trans;          // [R][4] host libm: internalM12M1, externalM12M1, lnSum(ext, int), lnSum(ext, M12D)
    const double *model_mean;     // pore_model means [4^9]
    double ln_c, c, two_sigma2;   // log(1/sqrt(2 sigma^2 pi)), that factor, 2 sigma^2 (sigma = 0.14), host libm
    double inv_two_sigma2;        // 1 / (2 sigma^2), correctly rounded
    double d2d, d2m, i2m, m2d, m2i, i2i;   // eln() of HMM_TransitionProbs_DNA_R10 (src/config.h:42), host libm
    const uint64_t *rec_off;      // [R+1] record capacity offsets
    dnb_eventalign_rec *recs;
    uint32_t *n_rec;              // [R]
    int *status;                  // [R] in: DNB_READ_OK or a host-side rejection; out: DNB_READ_*
    unsigned int *next_read;      // work counter (zeroed before the launch)
    const uint32_t *order;        // optional processing order (longest first): slot -> read; nullptr = identity
    double *scratch_obs;          // [warps][t_max]
    uint32_t *scratch_ev;         // [warps][t_max]
    uint8_t *scratch_bt;          // [warps][t_max][96] backtrace codes: I (2 bits) | M (2 bits) << 2 | D (1 bit) << 4
};
unsigned dnb_eventalign_grid(int device);
unsigned dnb_eventalign_warps_per_block(void);
size_t dnb_eventalign_bt_row_bytes(void);
void dnb_launch_eventalign(const DnbEaArgs &a, unsigned grid, cudaStream_t s);
// window-parallel form of the same stage (eventalign_wp.cu; the default): tot_ref / tot_align = total reference bases /
// aligned events of the batch; its workspace comes from `alloc` (caller-owned); waits once per round for two counters
// (on sync_ev, a cudaEventBlockingSync event, when given)
struct DnbAlloc {
    void *(*fn)(void *user, size_t bytes);
    void *user;
};
cudaError_t dnb_run_eventalign_wp(const DnbEaArgs &a, uint64_t tot_ref, uint64_t tot_align, int device, cudaStream_t s,
                                  const DnbAlloc &alloc, cudaEvent_t sync_ev);

// ---- DNN input tensors (features.cu): SURVEY s.8 row f2, consumes eventalign's records on the device --------------
struct DnbFeatArgs {
    uint32_t n_reads;
    const uint64_t *rec_off;      // as DnbEaArgs
    const dnb_eventalign_rec *recs;
    const uint32_t *n_rec;
    int *status;                  // [R] in: eventalign's status; out: UNDEFINED (non-monotone records) / OVERFLOW (row capacity)
    const uint64_t *ref_off;
    const char *ref;
    const int32_t *r2q;
    const uint64_t *raw_off;      // [R] first sample of the read in raw_f32 / raw_i16 (whichever raw_kind says)
    const uint8_t *raw_kind;      // [R] 0 = float32 pA, 1 = int16 DAC
    const float *raw_f32;
    const int16_t *raw_i16;
    const float *dac_offset, *dac_scale;   // [R], read only where raw_kind is 1
    const uint64_t *ev_off;       // [R+1]; read r's n_events + 1 event starts begin at ev_start[ev_off[r] + r]
    const uint32_t *ev_start;
    const double *shift, *scale;  // [R] r.scalings
    const uint8_t *is_reverse;    // [R]
    const uint32_t *ref_start, *ref_end;   // [R] r.refStart / r.refEnd
    const uint64_t *called_off;   // [R+1]
    const uint32_t *called;       // sorted keys of r.refCoordToCalls
    const uint64_t *pos_off;      // [R+1] row capacity offsets
    float *signal;                // [pos_off[R]][DNB_RAWDEPTH]
    float *core, *residual;
    uint32_t *coords, *ref_index, *query_index;
    int32_t *quality;
    uint32_t *n_pos;              // [R]
    unsigned int *next_read;      // work counter (zeroed before the launch)
    const uint32_t *order;        // optional processing order (longest first): slot -> read; nullptr = identity
};
unsigned dnb_features_grid(int device);
void dnb_launch_features(const DnbFeatArgs &a, unsigned grid, cudaStream_t s);
