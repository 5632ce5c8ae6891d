// TEST INFRASTRUCTURE -- oracle/_ref harness.  Never shipped, never on the product path.
//
// A flat C interface (ctypes-friendly) over the UNMODIFIED reference translation units of
// MBoemo/DNAscent, compiled where they lie under /root/reference by oracle/Makefile:
//   src/scrappie/event_detection.c  (detect_events)
//   src/event_handling.cpp          (normaliseEvents and its non-static helpers)
//   src/probability.cpp             (eexp/eln/lnSum/lnProd/lnGreaterThan/*PDF)
//   src/detect.cpp                  (sequenceProbability, getPOIs, llAcrossRead)
//   src/data_IO.cpp, src/htsInterface.cpp, src/common.cpp, src/alignment.cpp (link closure)
// Reads are built through the reference's own DNAscent::read constructor (src/reads.h:210) from a
// hand-encoded bam1_t, so parseCigar / getQuerySequence / reverseComplement are the reference's.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// the resulting library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <string>
#include <vector>
#include <omp.h>

#include "reads.h"
#include "config.h"
#include "event_handling.h"
#include "probability.h"
#include "detect.h"
#include "alignment.h"
#include "data_IO.h"
#include "scrappie/event_detection.h"

// normally defined in src/main/DNAscent.cpp:61
Global_Config Pore_Substrate_Config;

// non-static in src/alignment.cpp:193, not declared in its header
std::pair<double, std::vector<std::string>> builtinViterbi(std::vector<double> &observations, std::string &sequence,
                                                           PoreParameters scalings, bool flip);
// non-static helpers of src/event_handling.cpp (external linkage, not in its header)
PoreParameters estimateScaling_quantiles(std::vector<double> &signal_means, std::string &sequence,
                                         std::vector<unsigned int> &kmer_ranks, bool useFitPoreModel);
PoreParameters estimateScaling_theilSen(std::vector<double> &signals, std::vector<unsigned int> &kmer_ranks,
                                        PoreParameters s, bool useFitPoreModel);
std::pair<std::vector<double>, std::vector<unsigned int>> adaptive_banded_simple_event_align(
    DNAscent::read &r, std::vector<unsigned int> &kmer_ranks_query, std::vector<unsigned int> &kmer_ranks_ref,
    bool useFitPoreModel);

namespace {

std::map<std::string, std::string> g_reference;
std::map<std::string, IndexEntry> g_index;
bam_hdr_t *g_hdr = nullptr;

// layout mirror of BandedAlignQCs (src/reads.h:34-53; its data members are private)
struct QCMirror {
    double avg_log_emission;
    bool spanned, set;
    unsigned int maxGap;
};
static_assert(sizeof(QCMirror) == sizeof(BandedAlignQCs), "BandedAlignQCs layout changed");

struct Handle {
    DNAscent::read *r = nullptr;
    // stage captures (filled by dnbref_normalise_staged)
    PoreParameters rough{};
    std::vector<double> cleanedSignals;
    std::vector<unsigned int> cleanedRanks;
    size_t et_n = 0;
};

std::vector<std::pair<double, double>> &table(int which) {
    switch (which) {
        case 0: return Pore_Substrate_Config.pore_model;
        case 1: return Pore_Substrate_Config.unlabelled_model;
        default: return Pore_Substrate_Config.analogue_model;
    }
}

void configure_constants() {
    // the non-file half of Global_Config::configure_DNA_R10 (src/config.h:44-63)
    Pore_Substrate_Config.kmer_len = 9;
    Pore_Substrate_Config.windowLength_align = 50;
    Pore_Substrate_Config.HMM_config = Pore_Substrate_Config.HMM_TransitionProbs_DNA_R10;
    Pore_Substrate_Config.AdaptiveBanded_config = Pore_Substrate_Config.AdaptiveBanded_Params_DNA_R10;
}

}  // namespace

extern "C" {

// ---- configuration -------------------------------------------------------------------------
// Real loader: parses /root/reference/pore_models/*.model (container only).
int dnbref_configure_from_files(void) {
    try {
        Pore_Substrate_Config.configure_DNA_R10();
    } catch (...) {
        return -1;
    }
    return 0;
}

// One table through the reference's own file parser (src/data_IO.cpp:144-242): e.g. analogue_model re-pointed at
// r10.4.1_EdU_gaussian.model, which src/config.h:54 never loads (SURVEY s.0.2).  Container only.
int dnbref_load_model_file(int which, const char *filename, int fit_stdv) {
    configure_constants();
    try {
        table(which) = fit_stdv ? import_poreModel_fitStdv(filename, Pore_Substrate_Config.kmer_len)
                                : import_poreModel_staticStdv(filename, Pore_Substrate_Config.kmer_len);
    } catch (...) {
        return -1;
    }
    return 0;
}

// GPU box (no /root/reference): tables come from the committed fixtures.
int dnbref_set_model(int which, const double *mean, const double *stdv, size_t n) {
    configure_constants();
    auto &t = table(which);
    t.assign(n, std::make_pair(0.0, 0.0));
    for (size_t i = 0; i < n; i++) t[i] = std::make_pair(mean[i], stdv[i]);
    return 0;
}

size_t dnbref_get_model(int which, double *mean, double *stdv, size_t cap) {
    auto &t = table(which);
    for (size_t i = 0; i < t.size() && i < cap; i++) {
        mean[i] = t[i].first;
        stdv[i] = t[i].second;
    }
    return t.size();
}

unsigned int dnbref_kmer2index(const char *kmer, unsigned int k) {
    std::string s(kmer);
    return kmer2index(s, k);
}

int dnbref_set_reference(const char *name, const char *seq, size_t len) {
    g_reference[std::string(name)] = std::string(seq, len);
    if (!g_hdr) {
        g_hdr = (bam_hdr_t *)calloc(1, sizeof(bam_hdr_t));
        g_hdr->n_targets = 1;
        g_hdr->target_name = (char **)calloc(1, sizeof(char *));
        g_hdr->target_len = (uint32_t *)calloc(1, sizeof(uint32_t));
    }
    free(g_hdr->target_name[0]);
    g_hdr->target_name[0] = strdup(name);
    g_hdr->target_len[0] = (uint32_t)len;
    return 0;
}

// ---- reads ---------------------------------------------------------------------------------
// seq_bam: the SEQ column of the SAM record (reference-strand orientation), cigar: BAM-encoded ops.
void *dnbref_read_new(const char *qname, const char *seq_bam, uint32_t l_seq, const uint32_t *cigar,
                      uint32_t n_cigar, int flag, int32_t pos, const float *raw_pA, size_t n_raw) {
    if (!g_hdr) return nullptr;
    bam1_t *b = (bam1_t *)calloc(1, sizeof(bam1_t));
    size_t lq = strlen(qname) + 1;
    size_t extranul = (4 - (lq & 3)) & 3;
    size_t l_qname = lq + extranul;
    size_t l_data = l_qname + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
    b->data = (uint8_t *)calloc(l_data, 1);
    b->l_data = (int)l_data;
    b->m_data = (uint32_t)l_data;
    memcpy(b->data, qname, lq - 1);
    memcpy(b->data + l_qname, cigar, 4 * (size_t)n_cigar);
    uint8_t *s = b->data + l_qname + 4 * (size_t)n_cigar;
    for (uint32_t i = 0; i < l_seq; i++) {
        uint8_t code;
        switch (seq_bam[i]) {
            case 'A': code = 1; break;
            case 'C': code = 2; break;
            case 'G': code = 4; break;
            case 'T': code = 8; break;
            default: code = 15; break;
        }
        s[i >> 1] |= (uint8_t)(code << ((~i & 1) << 2));
    }
    memset(s + (l_seq + 1) / 2, 0xff, l_seq);
    b->core.tid = 0;
    b->core.pos = pos;
    b->core.qual = 60;
    b->core.l_qname = (uint8_t)l_qname;
    b->core.l_extranul = (uint8_t)extranul;
    b->core.flag = (uint16_t)flag;
    b->core.n_cigar = n_cigar;
    b->core.l_qseq = (int32_t)l_seq;
    b->core.mtid = -1;
    b->core.mpos = -1;

    Handle *h = new Handle();
    try {
        h->r = new DNAscent::read(b, g_hdr, g_index, g_reference);  // src/reads.h:210
    } catch (...) {
        delete h;
        return nullptr;
    }
    h->r->raw.resize(n_raw);
    for (size_t i = 0; i < n_raw; i++) h->r->raw[i] = (double)raw_pA[i];  // float32-exact, cf. src/pod5.cpp:60
    return h;
}

void dnbref_read_free(void *hv) {
    Handle *h = (Handle *)hv;
    if (!h) return;
    delete h->r;
    delete h;
}

// hot-path inputs exactly as normaliseEvents sees them (sequencing orientation)
size_t dnbref_read_basecall(void *hv, char *out, size_t cap) {
    const std::string &s = ((Handle *)hv)->r->basecall;
    if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
    return s.size();
}
size_t dnbref_read_refseq(void *hv, char *out, size_t cap) {
    const std::string &s = ((Handle *)hv)->r->referenceSeqMappedTo;
    if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
    return s.size();
}
// dense queryToRef: out[q] = ref index or -1 if the map has no entry (src/reads.h:193)
size_t dnbref_read_query_to_ref(void *hv, int32_t *out, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    size_t n = r->basecall.size();
    for (size_t q = 0; q < n && q < cap; q++) {
        auto it = r->queryToRef.find((unsigned)q);
        out[q] = it == r->queryToRef.end() ? -1 : (int32_t)it->second;
    }
    return n;
}
size_t dnbref_read_ref_to_query(void *hv, int32_t *out, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    size_t n = r->referenceSeqMappedTo.size();
    for (size_t q = 0; q < n && q < cap; q++) {
        auto it = r->refToQuery.find((unsigned)q);
        out[q] = it == r->refToQuery.end() ? -1 : (int32_t)it->second;
    }
    return n;
}
int dnbref_read_is_reverse(void *hv) { return ((Handle *)hv)->r->isReverse ? 1 : 0; }
int dnbref_read_ref_start(void *hv) { return ((Handle *)hv)->r->refStart; }
int dnbref_read_ref_end(void *hv) { return ((Handle *)hv)->r->refEnd; }

// ---- the hot path --------------------------------------------------------------------------
// The reference entry point, untouched (src/event_handling.cpp:544).
int dnbref_normalise(void *hv, int useFit) {
    Handle *h = (Handle *)hv;
    normaliseEvents(*h->r, useFit != 0);
    return 0;
}

#ifndef DNB_SHIM_BUILD
// normaliseEvents, then re-derivation of the intermediate stages through the reference's own helper
// functions so that per-stage goldens exist: rough (quantile) scaling, cleaned (signal,rank) vectors.
// Returns 0 if the re-derived alignment equals the one normaliseEvents produced, 1 otherwise.
int dnbref_normalise_staged(void *hv, int useFitI) {
    Handle *h = (Handle *)hv;
    DNAscent::read &r = *h->r;
    bool useFit = useFitI != 0;
    normaliseEvents(r, useFit);

    size_t k = Pore_Substrate_Config.kmer_len;
    std::vector<double> event_means;
    for (auto &e : r.events) event_means.push_back(e.mean);
    std::vector<unsigned int> rq(r.basecall.size() - k + 1), rr(r.referenceSeqMappedTo.size() - k + 1);
    for (size_t i = 0; i < rq.size(); i++) { std::string km = r.basecall.substr(i, k); rq[i] = kmer2index(km, k); }
    for (size_t i = 0; i < rr.size(); i++) { std::string km = r.referenceSeqMappedTo.substr(i, k); rr[i] = kmer2index(km, k); }

    h->rough = estimateScaling_quantiles(event_means, r.referenceSeqMappedTo, rr, useFit);

    auto savedAlign = r.eventAlignment;
    PoreParameters savedScalings = r.scalings;
    BandedAlignQCs savedQC = r.alignmentQCs;
    r.eventAlignment.clear();
    r.scalings = h->rough;
    auto seg = adaptive_banded_simple_event_align(r, rq, rr, useFit);
    h->cleanedSignals = seg.first;
    h->cleanedRanks = seg.second;
    PoreParameters ts = estimateScaling_theilSen(seg.first, seg.second, r.scalings, useFit);
    bool same = true;
    if (ts.shift == -1.) r.eventAlignment.clear();
    if (r.eventAlignment != savedAlign) same = false;
    if (ts.shift != savedScalings.shift || ts.scale != savedScalings.scale) same = false;
    if (memcmp(&savedQC, &r.alignmentQCs, sizeof(QCMirror)) != 0) same = false;
    r.eventAlignment = savedAlign;
    r.scalings = savedScalings;
    r.alignmentQCs = savedQC;
    return same ? 0 : 1;
}

#endif  // !DNB_SHIM_BUILD

size_t dnbref_n_events(void *hv) { return ((Handle *)hv)->r->events.size(); }
// means[j] = r.events[j].mean ; raw_len[j] = r.events[j].raw.size()
size_t dnbref_events(void *hv, double *means, uint32_t *raw_len, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    for (size_t j = 0; j < r->events.size() && j < cap; j++) {
        if (means) means[j] = r->events[j].mean;
        if (raw_len) raw_len[j] = (uint32_t)r->events[j].raw.size();
    }
    return r->events.size();
}
// concatenation of r.events[j].raw (what src/alignment.cpp:704 consumes)
size_t dnbref_events_raw_concat(void *hv, double *out, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    size_t n = 0;
    for (auto &e : r->events)
        for (double v : e.raw) {
            if (out && n < cap) out[n] = v;
            n++;
        }
    return n;
}
size_t dnbref_alignment(void *hv, uint32_t *ev, uint32_t *km, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    for (size_t j = 0; j < r->eventAlignment.size() && j < cap; j++) {
        ev[j] = r->eventAlignment[j].first;
        km[j] = r->eventAlignment[j].second;
    }
    return r->eventAlignment.size();
}
// out = {shift, scale, eventsPerBase, rough_shift, rough_scale, avg_log_emission, spanned, maxGap, qc_set}
void dnbref_scalars(void *hv, double *out) {
    Handle *h = (Handle *)hv;
    QCMirror q;
    memcpy(&q, &h->r->alignmentQCs, sizeof(q));
    out[0] = h->r->scalings.shift;
    out[1] = h->r->scalings.scale;
    out[2] = h->r->scalings.eventsPerBase;
    out[3] = h->rough.shift;
    out[4] = h->rough.scale;
    out[5] = q.avg_log_emission;
    out[6] = q.spanned ? 1.0 : 0.0;
    out[7] = (double)q.maxGap;
    out[8] = q.set ? 1.0 : 0.0;
}
size_t dnbref_cleaned(void *hv, double *sig, uint32_t *ranks, size_t cap) {
    Handle *h = (Handle *)hv;
    for (size_t j = 0; j < h->cleanedSignals.size() && j < cap; j++) {
        sig[j] = h->cleanedSignals[j];
        ranks[j] = h->cleanedRanks[j];
    }
    return h->cleanedSignals.size();
}

// scrappie segmentation alone (src/scrappie/event_detection.c:268). Returns et.n.
size_t dnbref_detect_events(const float *raw_pA, size_t n, uint64_t *start, float *length, float *mean, float *stdv,
                            size_t cap) {
    std::vector<double> raw(n);
    for (size_t i = 0; i < n; i++) raw[i] = (double)raw_pA[i];
    event_table et = detect_events(raw.data(), n, event_detection_defaults);
    for (size_t i = 0; i < et.n && i < cap; i++) {
        if (start) start[i] = et.event[i].start;
        if (length) length[i] = et.event[i].length;
        if (mean) mean[i] = et.event[i].mean;
        if (stdv) stdv[i] = et.event[i].stdv;
    }
    size_t n_ev = et.n;
    free(et.event);
    return n_ev;
}

#ifndef DNB_SHIM_BUILD
// Theil-Sen alone (src/event_handling.cpp:24)
void dnbref_theil_sen(const double *sig, const uint32_t *ranks, size_t n, double shift, double scale, int useFit,
                      double *out_shift, double *out_scale) {
    std::vector<double> s(sig, sig + n);
    std::vector<unsigned int> rk(ranks, ranks + n);
    PoreParameters p;
    p.shift = shift;
    p.scale = scale;
    PoreParameters o = estimateScaling_theilSen(s, rk, p, useFit != 0);
    *out_shift = o.shift;
    *out_scale = o.scale;
}

#endif  // !DNB_SHIM_BUILD

// ---- analogue likelihood path (src/detect.cpp:235-574) ---------------------------------------
double dnbref_sequence_probability(const double *obs, size_t n_obs, const char *seq, size_t windowSize, int useBrdU,
                                   double shift, double scale, double eventsPerBase, size_t BrdUStart, size_t BrdUEnd) {
    std::vector<double> o(obs, obs + n_obs);
    std::string s(seq);
    PoreParameters p;
    p.shift = shift;
    p.scale = scale;
    p.eventsPerBase = eventsPerBase;
    return sequenceProbability(o, s, windowSize, useBrdU != 0, p, BrdUStart, BrdUEnd);
}

// llAcrossRead on a read that has been through dnbref_normalise; returns #calls, fills (global ref pos, LLR)
size_t dnbref_ll_across_read(void *hv, unsigned int windowLength, int32_t *pos, double *llr, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    r->refCoordToCalls.clear();
    llAcrossRead(*r, windowLength);
    size_t n = 0;
    for (auto &kv : r->refCoordToCalls) {
        if (n < cap) {
            pos[n] = (int32_t)kv.first;
            llr[n] = kv.second.first;
        }
        n++;
    }
    return n;
}

// ---- eventalign / builtinViterbi (src/alignment.cpp:193-516, 547-744): SURVEY s.8 row f1 -------------------
#ifndef DNB_SHIM_BUILD
// one window: returns the number of states on the path; state i is (index[i], type[i]) with type 0=D 1=M 2=I
size_t dnbref_builtin_viterbi(const double *obs, size_t n_obs, const char *seq, double shift, double scale,
                              double eventsPerBase, double *score, int32_t *index, uint8_t *type, size_t cap) {
    std::vector<double> o(obs, obs + n_obs);
    std::string s(seq);
    PoreParameters p;
    p.shift = shift;
    p.scale = scale;
    p.eventsPerBase = eventsPerBase;
    std::pair<double, std::vector<std::string>> res = builtinViterbi(o, s, p, false);
    *score = res.first;
    size_t n = 0;
    for (const std::string &lab : res.second) {
        if (n < cap) {
            index[n] = std::stoi(lab.substr(0, lab.find('_')));
            const char t = lab[lab.find('_') + 1];
            type[n] = t == 'D' ? 0 : t == 'M' ? 1 : 2;
        }
        n++;
    }
    return n;
}
#endif  // !DNB_SHIM_BUILD

// eventalign on a read that has been through normaliseEvents (the reference's own in the ref build, the shim's in the
// shim build); returns strlen(humanReadable_eventalignOut) and copies up to cap bytes
size_t dnbref_eventalign(void *hv, unsigned int windowLength, char *out, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    r->refCoordToAP.clear();
    r->humanReadable_eventalignOut.clear();
    eventalign(*r, windowLength);
    const std::string &s = r->humanReadable_eventalignOut;
    if (out && cap) memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
    return s.size();
}

// what eventalign left in r.refCoordToAP through r.addSignal (reads.h:288-372), as the DNN input builders see it:
// signal tensor [P*RAWDEPTH], core / residual k-mer indices [P], reference coordinates [P]; returns P
size_t dnbref_aligned_positions(void *hv, float *signal, float *core, float *residual, uint32_t *coords, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    const size_t P = r->refCoordToAP.size();
    if (P == 0 || P > cap) return P;
    std::vector<float> sg = r->makeSignalTensor(), co = r->makeCoreSequenceTensor(), re = r->makeResidualSequenceTensor();
    std::vector<unsigned int> rc = r->getReferenceCoords();
    memcpy(signal, sg.data(), sg.size() * sizeof(float));
    memcpy(core, co.data(), co.size() * sizeof(float));
    memcpy(residual, re.data(), re.size() * sizeof(float));
    for (size_t i = 0; i < rc.size(); i++) coords[i] = rc[i];
    return P;
}
size_t dnbref_rawdepth(void) { return RAWDEPTH; }
// the per-position bookkeeping runCNN reads next to the tensors (detect.cpp:668-671), in tensor order:
// getReferenceIndices, getQueryIndices and each AlignedPosition's alignment quality (the window's indelScore)
size_t dnbref_aligned_indices(void *hv, uint32_t *ref_index, uint32_t *query_index, int32_t *quality, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    const size_t P = r->refCoordToAP.size();
    if (P == 0 || P > cap) return P;
    std::vector<unsigned int> ri = r->getReferenceIndices(), qi = r->getQueryIndices();
    for (size_t i = 0; i < P; i++) { ref_index[i] = ri[i]; query_index[i] = qi[i]; }
    size_t o = 0;
    if (r->strand == "fwd") for (auto p = r->refCoordToAP.begin(); p != r->refCoordToAP.end(); p++) quality[o++] = (int32_t)p->second->getAlignmentQuality();
    else for (auto p = r->refCoordToAP.rbegin(); p != r->refCoordToAP.rend(); p++) quality[o++] = (int32_t)p->second->getAlignmentQuality();
    return P;
}

// ---- probability.cpp helpers -----------------------------------------------------------------
double dnbref_eexp(double x) { return eexp(x); }
// returns 1 and leaves *out untouched when the reference throws NegativeLog
int dnbref_eln(double x, double *out) {
    try {
        *out = eln(x);
    } catch (NegativeLog &) {
        return 1;
    }
    return 0;
}
double dnbref_lnSum(double a, double b) { return lnSum(a, b); }
double dnbref_lnProd(double a, double b) { return lnProd(a, b); }
int dnbref_lnGreaterThan(double a, double b) { return lnGreaterThan(a, b) ? 1 : 0; }
double dnbref_uniformPDF(double lb, double ub, double x) { return uniformPDF(lb, ub, x); }
double dnbref_normalPDF(double mu, double sigma, double x) { return normalPDF(mu, sigma, x); }
double dnbref_cauchyPDF(double loc, double scale, double x) { return cauchyPDF(loc, scale, x); }

// ---- CPU baseline: the read loop of src/detect.cpp:852-876 minus I/O and DNN -------------------
// Returns wall seconds; *failed = reads whose eventAlignment came back empty (detect.cpp:879).
double dnbref_bench_normalise(void **handles, size_t n, int threads, int useFit, int *failed) {
    int nfail = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(+ : nfail)
    for (size_t i = 0; i < n; i++) {
        Handle *h = (Handle *)handles[i];
        normaliseEvents(*h->r, useFit != 0);
        if (h->r->eventAlignment.empty()) nfail++;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (failed) *failed = nfail;
    return std::chrono::duration<double>(t1 - t0).count();
}

// CPU baseline of the analogue path (row a15): detect --HMM's loop body, detect.cpp:876-885 = normaliseEvents + llAcrossRead.
// *calls = LLR calls made over all reads (refCoordToCalls entries); returns wall seconds.
double dnbref_bench_hmm(void **handles, size_t n, int threads, unsigned int windowLength, int *failed, long long *calls) {
    int nfail = 0;
    long long ncalls = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(+ : nfail, ncalls)
    for (size_t i = 0; i < n; i++) {
        Handle *h = (Handle *)handles[i];
        normaliseEvents(*h->r, false);
        if (h->r->eventAlignment.empty()) { nfail++; continue; }
        h->r->refCoordToCalls.clear();
        llAcrossRead(*h->r, windowLength);
        ncalls += (long long)h->r->refCoordToCalls.size();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (failed) *failed = nfail;
    if (calls) *calls = ncalls;
    return std::chrono::duration<double>(t1 - t0).count();
}

// CPU baseline of the chain (rows f1-f2): detect.cpp:876-888 per read = normaliseEvents + eventalign (whose
// r.addSignal calls are what the tensor builders read); reads that fail normalisation skip eventalign (detect.cpp:879)
double dnbref_bench_chain(void **handles, size_t n, int threads, unsigned int windowLength, int *failed) {
    int nfail = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(+ : nfail)
    for (size_t i = 0; i < n; i++) {
        Handle *h = (Handle *)handles[i];
        normaliseEvents(*h->r, false);
        if (h->r->eventAlignment.empty()) { nfail++; continue; }
        h->r->refCoordToAP.clear();
        eventalign(*h->r, windowLength);
        std::vector<float> sg = h->r->makeSignalTensor(), co = h->r->makeCoreSequenceTensor(), re = h->r->makeResidualSequenceTensor();
        if (sg.size() != co.size() * RAWDEPTH || re.size() != co.size()) nfail++;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (failed) *failed = nfail;
    return std::chrono::duration<double>(t1 - t0).count();
}

#ifdef DNB_SHIM_BUILD
}  // extern "C"
#include "dnascent_shim.h"
extern "C" {
// shim build only: the batched entry points the patched read loop of detect.cpp would call
int dnbshim_normalise_batch(void **handles, size_t n) {
    std::vector<DNAscent::read *> reads(n);
    for (size_t i = 0; i < n; i++) reads[i] = ((Handle *)handles[i])->r;
    dnb_shim::normaliseEvents_batch(reads, false);
    return 0;
}
int dnbshim_ll_across_read_batch(void **handles, size_t n, unsigned int windowLength) {
    std::vector<DNAscent::read *> reads(n);
    for (size_t i = 0; i < n; i++) {
        reads[i] = ((Handle *)handles[i])->r;
        reads[i]->refCoordToCalls.clear();
    }
    dnb_shim::llAcrossRead_batch(reads, windowLength);
    return 0;
}
// detect --HMM as one resident chain: normaliseEvents + llAcrossRead on the device (dnb_submit_llr)
int dnbshim_normalise_ll_batch(void **handles, size_t n, unsigned int windowLength) {
    std::vector<DNAscent::read *> reads(n);
    for (size_t i = 0; i < n; i++) {
        reads[i] = ((Handle *)handles[i])->r;
        reads[i]->refCoordToCalls.clear();
    }
    dnb_shim::normalise_llAcrossRead_batch(reads, windowLength);
    return 0;
}
// calls recorded by (batched) llAcrossRead, as dnbref_ll_across_read returns them
size_t dnbshim_calls(void *hv, int32_t *pos, double *llr, size_t cap) {
    DNAscent::read *r = ((Handle *)hv)->r;
    size_t n = 0;
    for (auto &kv : r->refCoordToCalls) {
        if (n < cap) { pos[n] = (int32_t)kv.first; llr[n] = kv.second.first; }
        n++;
    }
    return n;
}
// eventalign + DNN input tensors built on the device (row f2); results kept until the next call, read back per read
static std::vector<dnb_shim::DnnInputs> g_dnn_inputs;
int dnbshim_eventalign_features_batch(void **handles, size_t n, unsigned int windowLength) {
    std::vector<DNAscent::read *> reads(n);
    for (size_t i = 0; i < n; i++) reads[i] = ((Handle *)handles[i])->r;
    try {
        dnb_shim::eventalign_features_batch(reads, windowLength, g_dnn_inputs);
    } catch (NegativeLog &) {
        return 1;
    }
    return 0;
}
// the resident chain: normaliseEvents + eventalign + tensors in one device-resident batch
int dnbshim_normalise_eventalign_batch(void **handles, size_t n, unsigned int windowLength) {
    std::vector<DNAscent::read *> reads(n);
    for (size_t i = 0; i < n; i++) reads[i] = ((Handle *)handles[i])->r;
    try {
        dnb_shim::normalise_eventalign_batch(reads, windowLength, g_dnn_inputs);
    } catch (NegativeLog &) {
        return 1;
    }
    return 0;
}
size_t dnbshim_dnn_inputs(size_t i, float *signal, float *core, float *residual, uint32_t *coords, uint32_t *ref_index,
                          uint32_t *query_index, int32_t *quality, size_t cap) {
    if (i >= g_dnn_inputs.size()) return 0;
    const dnb_shim::DnnInputs &d = g_dnn_inputs[i];
    const size_t P = d.core.size();
    if (P == 0 || P > cap) return P;
    memcpy(signal, d.signal.data(), d.signal.size() * sizeof(float));
    memcpy(core, d.core.data(), P * sizeof(float));
    memcpy(residual, d.residual.data(), P * sizeof(float));
    for (size_t j = 0; j < P; j++) {
        coords[j] = d.refCoords[j]; ref_index[j] = d.refIndices[j]; query_index[j] = d.queryIndices[j];
        quality[j] = d.alignmentQuality[j];
    }
    return P;
}
void dnbshim_shutdown(void) { dnb_shim::shutdown(); }
// one process, several GPUs: call before the first batched call (after dnbshim_shutdown when a context exists)
void dnbshim_set_devices(const int *devices, int n) { dnb_shim::set_devices(std::vector<int>(devices, devices + n)); }
unsigned long dnbshim_batches_on_device(int device) { return dnb_shim::batches_on_device(device); }
#endif  // DNB_SHIM_BUILD

int dnbref_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
