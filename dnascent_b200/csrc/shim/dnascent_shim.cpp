// dnascent_shim.cpp -- the reference's C++ symbols for the detect signal hot path, on top of libdnascent_b200.
// See dnascent_shim.h for what it replaces and INTEGRATION.md for how it is linked into the reference.
//
// This file is compiled against the REFERENCE's headers (reads.h, config.h, probability.h, event_detection.h):
// the data contract is DNAscent::read itself (src/reads.h:178-208).  All arithmetic happens on the GPU behind the
// C ABI; this layer only marshals:  r.raw (double, float32-exact: src/pod5.cpp:60) -> float32 staging,
// r.queryToRef (std::map) -> dense int32, and the results back into r.events / r.eventAlignment / r.scalings /
// r.alignmentQCs.  There is no CPU fallback: any library error aborts with the library's message, like the
// reference's own `assert(et.n > 0)` (src/event_handling.cpp:547).
#include "dnascent_shim.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "config.h"
#include "event_handling.h"
#include "alignment.h"
#include "common.h"
#include "probability.h"
#include "scrappie/event_detection.h"
#include "dnascent_b200.h"

namespace {

std::mutex g_mu;
dnb_ctx *g_ctx = nullptr;
int g_device = -1;
std::vector<int> g_devices;   // more than one entry: one context (one process) drives all of them
std::atomic<unsigned long> g_batches_on[DNB_MAX_DEVICES];   // submissions each device ran (diagnostics)

void note_device(dnb_batch *b) {
    const int dev = dnb_batch_device(b);
    if (dev >= 0 && dev < DNB_MAX_DEVICES) g_batches_on[dev]++;
}

[[noreturn]] void die(const char *where, int rc) {
    std::fprintf(stderr, "dnascent_b200 shim: %s failed: %s (%s)\n", where, dnb_strerror(rc), dnb_last_error());
    std::abort();
}

void load_table(dnb_ctx *ctx, int which, const std::vector<std::pair<double, double>> &t) {
    if (t.empty()) return;   // table not configured (e.g. a tool that never scores analogues)
    std::vector<double> mean(t.size()), stdv(t.size());
    for (size_t i = 0; i < t.size(); i++) {
        mean[i] = t[i].first;
        stdv[i] = t[i].second;
    }
    int rc = dnb_load_model(ctx, which, mean.data(), stdv.data(), t.size());
    if (rc != DNB_OK) die("dnb_load_model", rc);
}

// r.raw is vector<double> but every value is a float (pod5.cpp:60, fast5.cpp:103-107): narrowing is lossless
void narrow_signal(const std::vector<double> &raw, std::vector<float> &out) {
    out.resize(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        const float f = (float)raw[i];
        if ((double)f != raw[i]) {
            std::fprintf(stderr, "dnascent_b200 shim: raw[%zu] = %.17g is not float32-exact\n", i, raw[i]);
            std::abort();
        }
        out[i] = f;
    }
}

struct Staged {
    std::vector<float> raw;
    std::vector<dnb_q2r_run> runs;
};

// r.queryToRef (std::map, keys ascending) as runs: consecutive query positions whose reference positions advance by
// one (aligned bases) or stay (the entries parseCigar gives insertions and soft clips, htsInterface.cpp:143-152).
// A `{L}M` read is one run: 16 bytes cross PCIe instead of 4 per base.
void q2r_runs(const std::map<unsigned int, unsigned int> &m, size_t query_len, std::vector<dnb_q2r_run> &out) {
    out.clear();
    for (const auto &kv : m) {
        if (kv.first >= query_len) continue;
        if (!out.empty()) {
            dnb_q2r_run &t = out.back();
            if (kv.first == t.q_start + t.len) {
                const int32_t step = (int32_t)kv.second - (t.r_start + t.stride * (int32_t)(t.len - 1));
                if (t.len == 1 && (step == 0 || step == 1)) { t.stride = step; t.len = 2; continue; }
                if (t.len > 1 && step == t.stride) { t.len++; continue; }
            }
        }
        out.push_back(dnb_q2r_run{kv.first, 1u, (int32_t)kv.second, 1});
    }
    if (out.empty()) out.reserve(1);   // a non-NULL pointer tells the library "runs given", even when there are none
}

void stage_read(const DNAscent::read &r, Staged &s, dnb_read_desc &d) {
    narrow_signal(r.raw, s.raw);
    q2r_runs(r.queryToRef, r.basecall.size(), s.runs);
    std::memset(&d, 0, sizeof(d));
    d.raw_pA = s.raw.data();
    d.n_samples = s.raw.size();
    d.query = r.basecall.data();
    d.query_len = (uint32_t)r.basecall.size();
    d.ref = r.referenceSeqMappedTo.data();
    d.ref_len = (uint32_t)r.referenceSeqMappedTo.size();
    d.query_to_ref = nullptr;
    d.q2r_runs = s.runs.data();
    d.n_q2r_runs = (uint32_t)s.runs.size();
}

// what src/event_handling.cpp:549-606 leaves in the read
void unpack_result(DNAscent::read &r, const dnb_read_result &o) {
    if (o.status == DNB_READ_UNDEFINED || o.status == DNB_READ_OVERFLOW) {
        // inputs on which the reference itself aborts or indexes out of range: fail the read the reference's way
        r.events.clear();
        r.eventAlignment.clear();
        return;
    }
    // the results come back in the compact wire format (1 B per event length, 2 bits per alignment step); they are
    // expanded on the fly, the alignment straight into the vector the reference keeps it in
    int rc;
    std::vector<uint32_t> start((size_t)o.n_events + 1);
    if ((rc = dnb_expand_events(&o, start.data())) != DNB_OK) die("dnb_expand_events", rc);
    r.events.clear();
    r.events.resize(o.n_events);
    for (uint32_t j = 0; j < o.n_events; j++) {
        event &e = r.events[j];
        e.mean = (double)o.event_mean[j];
        e.raw.assign(r.raw.begin() + start[j], r.raw.begin() + start[j + 1]);
    }
    static_assert(sizeof(std::pair<unsigned int, unsigned int>) == 2 * sizeof(uint32_t), "eventAlignment element layout");
    r.eventAlignment.assign(o.n_align, std::pair<unsigned int, unsigned int>(0u, 0u));
    if ((rc = dnb_expand_alignment(&o, reinterpret_cast<uint32_t *>(r.eventAlignment.data()))) != DNB_OK)
        die("dnb_expand_alignment", rc);
    r.alignmentQCs.recordQCs(o.avg_log_emission, o.spanned != 0, (unsigned int)o.max_gap);
    r.scalings.shift = o.shift;
    r.scalings.scale = o.scale;
    r.scalings.eventsPerBase = o.events_per_base;
}

}  // namespace

namespace dnb_shim {

void set_device(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_device = device;
    g_devices.clear();
}

void set_devices(const std::vector<int> &devices) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_devices = devices;
    if (!devices.empty()) g_device = devices[0];
}

dnb_ctx *context() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) return g_ctx;
    dnb_config cfg;
    dnb_default_config(&cfg);
    if (g_devices.empty()) {
        // DNB_DEVICES="0,1,2,3": the single `DNAscent detect` process (detect.cpp:852) feeds all of them
        if (const char *e = std::getenv("DNB_DEVICES"))
            for (const char *p = e; *p;) {
                char *end = nullptr;
                const long v = std::strtol(p, &end, 10);
                if (end == p) break;
                g_devices.push_back((int)v);
                p = (*end == ',') ? end + 1 : end;
            }
    }
    if (g_device < 0) {
        const char *e = std::getenv("DNB_DEVICE");
        g_device = !g_devices.empty() ? g_devices[0] : (e ? std::atoi(e) : 0);
    }
    cfg.device = g_device;
    if (g_devices.size() > 1) {
        cfg.n_devices = (int)std::min<size_t>(g_devices.size(), DNB_MAX_DEVICES);
        for (int k = 0; k < cfg.n_devices; k++) cfg.devices[k] = g_devices[k];
    }
    cfg.result_format = DNB_RESULT_COMPACT;
    cfg.min_average_log_emission = Pore_Substrate_Config.AdaptiveBanded_config.min_average_log_emission;
    cfg.max_gap_threshold = Pore_Substrate_Config.AdaptiveBanded_config.max_gap_threshold;
    cfg.bandwidth = Pore_Substrate_Config.AdaptiveBanded_config.bandwidth;
    dnb_ctx *ctx = nullptr;
    int rc = dnb_create(&ctx, &cfg);
    if (rc != DNB_OK) die("dnb_create", rc);
    load_table(ctx, DNB_MODEL_PORE, Pore_Substrate_Config.pore_model);
    load_table(ctx, DNB_MODEL_UNLABELLED, Pore_Substrate_Config.unlabelled_model);
    load_table(ctx, DNB_MODEL_ANALOGUE, Pore_Substrate_Config.analogue_model);
    g_ctx = ctx;
    return g_ctx;
}

void shutdown() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) dnb_destroy(g_ctx);
    g_ctx = nullptr;
}

unsigned long batches_on_device(int device) {
    return (device >= 0 && device < DNB_MAX_DEVICES) ? g_batches_on[device].load() : 0ul;
}

void normaliseEvents_batch(const std::vector<DNAscent::read *> &reads, bool useFitPoreModel) {
    if (reads.empty()) return;
    if (useFitPoreModel) {
        // every caller in the reference passes false (detect.cpp:875, alignment.cpp:855, trainCNN.cpp:318)
        std::fprintf(stderr, "dnascent_b200 shim: normaliseEvents(useFitPoreModel=true) is not supported\n");
        std::abort();
    }
    dnb_ctx *ctx = context();
    const size_t n = reads.size();
    std::vector<Staged> staged(n);
    std::vector<dnb_read_desc> descs(n);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) stage_read(*reads[i], staged[i], descs[i]);
    dnb_batch *b = nullptr;
    int rc = dnb_submit(ctx, descs.data(), n, &b);
    if (rc != DNB_OK) die("dnb_submit", rc);
    note_device(b);
    rc = dnb_wait(b);
    if (rc != DNB_OK) die("dnb_wait", rc);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) {
        dnb_read_result o;
        int rc2 = dnb_result(b, i, &o);
        if (rc2 != DNB_OK) die("dnb_result", rc2);
        unpack_result(*reads[i], o);
    }
    dnb_release(b);
}

}  // namespace dnb_shim

// ---- src/event_handling.h:13 ---------------------------------------------------------------------------------------
void normaliseEvents(DNAscent::read &r, bool useFitPoreModel) {
    std::vector<DNAscent::read *> one(1, &r);
    dnb_shim::normaliseEvents_batch(one, useFitPoreModel);
}

// ---- src/scrappie/event_detection.h:35 -------------------------------------------------------------------------------
extern "C" event_table detect_events(double *raw, size_t raw_size, detector_param const edparam) {
    const detector_param &d = event_detection_defaults;
    if (edparam.window_length1 != d.window_length1 || edparam.window_length2 != d.window_length2 ||
        edparam.threshold1 != d.threshold1 || edparam.threshold2 != d.threshold2 || edparam.peak_height != d.peak_height) {
        std::fprintf(stderr, "dnascent_b200 shim: detect_events supports event_detection_defaults only\n");
        std::abort();
    }
    event_table et = {0, 0, 0, nullptr};
    if (!raw || raw_size == 0) return et;
    std::vector<double> rv(raw, raw + raw_size);
    std::vector<float> f;
    narrow_signal(rv, f);
    // the caller frees et.event with free() (event_handling.cpp:575)
    dnb_event_t *ev = (dnb_event_t *)std::calloc(raw_size + 2, sizeof(dnb_event_t));
    size_t n = 0;
    int rc = dnb_detect_events(dnb_shim::context(), f.data(), raw_size, ev, raw_size + 2, &n);
    if (rc != DNB_OK) die("dnb_detect_events", rc);
    static_assert(sizeof(dnb_event_t) == sizeof(event_t), "event_t layout");
    et.n = n;
    et.start = 0;
    et.end = n;
    et.event = (event_t *)ev;
    return et;
}

// ---- src/probability.h:26-33 ------------------------------------------------------------------------------------------
double eexp(double x) { return dnb_eexp(x); }
double eln(double x) {
    double o = 0.0;
    if (dnb_eln(x, &o) == DNB_ERR_NEGATIVE_LOG) throw NegativeLog();   // src/probability.cpp:45
    return o;
}
double lnSum(double ln_x, double ln_y) { return dnb_lnSum(ln_x, ln_y); }
double lnProd(double ln_x, double ln_y) { return dnb_lnProd(ln_x, ln_y); }
bool lnGreaterThan(double ln_x, double ln_y) { return dnb_lnGreaterThan(ln_x, ln_y) != 0; }
double uniformPDF(double lb, double ub, double x) { return dnb_uniformPDF(lb, ub, x); }
double normalPDF(double mu, double sigma, double x) { return dnb_normalPDF(mu, sigma, x); }
double cauchyPDF(double location, double scale, double x) { return dnb_cauchyPDF(location, scale, x); }

#ifdef DNB_SHIM_WITH_HMM
// ---- src/detect.h:119,121 ------------------------------------------------------------------------------------------------
#include "detect.h"

namespace {

struct Site {
    size_t read;
    unsigned int posOnRef;
    std::vector<double> events;
};

// The event gathering of llAcrossRead for one read (src/detect.cpp:381-390, 399-510): which T positions are
// scored and with which events.  It uses the read's own refToQuery map with the reference's operators ([] inserts
// a zero for a missing key, at() throws), so gaps in the CIGAR behave identically.
void gather_sites(DNAscent::read &r, size_t read_idx, unsigned int w, std::vector<Site> &out) {
    const unsigned int k = Pore_Substrate_Config.kmer_len;
    const std::string &refSeq = r.referenceSeqMappedTo;
    std::vector<unsigned int> pois;
    for (unsigned int i = 2 * w; i < refSeq.length() - 2 * w; i++)
        if (refSeq[i] == 'T') pois.push_back(i);
    const auto &al = r.eventAlignment;
    size_t head = 0;
    if (r.isReverse) {
        head = al.size() - 1;
        std::reverse(pois.begin(), pois.end());
    }
    for (unsigned int p : pois) {
        (void)r.refToQuery.at(p);
        bool defined = true;
        for (size_t c = p - w; c < (size_t)p + w + k; c++) {
            const char ch = c < refSeq.size() ? refSeq[c] : 'N';
            if (ch != 'A' && ch != 'T' && ch != 'G' && ch != 'C') defined = false;
        }
        if (!defined || (size_t)p + w + k > refSeq.size()) continue;
        const unsigned int lo = r.refToQuery[p - w], hi = r.refToQuery[p + w];
        Site s;
        s.read = read_idx;
        s.posOnRef = p;
        bool first = true;
        auto consider = [&](size_t j) {
            if (lo <= al[j].second && al[j].second < hi) {
                if (first) { head = j; first = false; }
                const double ev = r.events[al[j].first].mean;
                if (ev > 0. && ev < 250.0) s.events.push_back(ev);
            }
        };
        if (r.isReverse) {
            for (long j = (long)head; j >= 0; j--) {
                consider((size_t)j);
                if (al[j].second < lo) { std::reverse(s.events.begin(), s.events.end()); break; }
            }
        } else {
            for (size_t j = head; j < al.size(); j++) {
                consider(j);
                if (al[j].second >= hi) break;
            }
        }
        if (s.events.size() < 2 * w - k) continue;
        out.push_back(std::move(s));
    }
}

}  // namespace

namespace dnb_shim {

void llAcrossRead_batch(const std::vector<DNAscent::read *> &reads, unsigned int w) {
    const unsigned int k = Pore_Substrate_Config.kmer_len;
    std::vector<Site> sites;
    for (size_t i = 0; i < reads.size(); i++) {
        DNAscent::read &r = *reads[i];
        r.humanReadable_detectOut = ">" + r.readID + " " + r.referenceMappedTo + " " + std::to_string(r.refStart) + " " +
                                    std::to_string(r.refEnd) + " " + (r.isReverse ? "rev" : "fwd") + "\n";
        if (!r.eventAlignment.empty()) gather_sites(r, i, w, sites);
    }
    const size_t n = sites.size();
    if (n == 0) return;
    const size_t snip = 2 * (size_t)w + k;
    std::vector<uint64_t> off(n + 1, 0);
    for (size_t s = 0; s < n; s++) off[s + 1] = off[s] + sites[s].events.size();
    std::vector<double> obs(off[n]), shift(n), scale(n), epb(n), la(n), lt(n);
    std::string seq(n * snip, 'A');
    for (size_t s = 0; s < n; s++) {
        const DNAscent::read &r = *reads[sites[s].read];
        std::copy(sites[s].events.begin(), sites[s].events.end(), obs.begin() + off[s]);
        std::memcpy(&seq[s * snip], r.referenceSeqMappedTo.data() + sites[s].posOnRef - w, snip);
        shift[s] = r.scalings.shift;
        scale[s] = r.scalings.scale;
        epb[s] = r.scalings.eventsPerBase;
    }
    int rc = dnb_sequence_probability_batch(context(), obs.data(), off.data(), seq.data(), shift.data(), scale.data(),
                                            epb.data(), n, w, la.data(), lt.data());
    if (rc != DNB_OK) die("dnb_sequence_probability_batch", rc);
    for (size_t s = 0; s < n; s++) {
        DNAscent::read &r = *reads[sites[s].read];
        const unsigned int p = sites[s].posOnRef, q = r.refToQuery.at(p);
        std::string kmerQuery = r.basecall.substr(q - k / 2, k), kmerRef = r.referenceSeqMappedTo.substr(p - k / 2, k);
        int globalPos = r.refStart + (int)p;
        if (r.isReverse) {
            globalPos = r.refEnd - (int)p - 1;
            kmerQuery = reverseComplement(kmerQuery);
            kmerRef = reverseComplement(kmerRef);
        }
        const double llr = la[s] - lt[s];   // detect.cpp:548
        r.humanReadable_detectOut += std::to_string(globalPos) + "\t" + std::to_string(llr) + "\t" + kmerRef + "\t" + kmerQuery + "\n";
        r.refCoordToCalls[globalPos] = std::make_pair(llr, 0.);
    }
}

}  // namespace dnb_shim

namespace dnb_shim {

// detect --HMM: normaliseEvents + llAcrossRead (detect.cpp:876-885) as one device-resident chain.  The T sites, the
// events of every site's window (readHead scan included) and both forward passes per site are computed on the GPU
// from the resident alignment (dnb_submit_llr); what stays here is the text of humanReadable_detectOut and
// r.refCoordToCalls (detect.cpp:527-572).
void normalise_llAcrossRead_batch(const std::vector<DNAscent::read *> &reads, unsigned int w) {
    const unsigned int k = Pore_Substrate_Config.kmer_len;
    const size_t n = reads.size();
    if (n == 0) return;
    dnb_ctx *ctx = context();
    std::vector<Staged> staged(n);
    std::vector<dnb_read_desc> descs(n);
    std::vector<dnb_read_extra> extra(n);
    std::vector<std::vector<int32_t>> r2q(n);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        stage_read(r, staged[i], descs[i]);
        const size_t rl = r.referenceSeqMappedTo.size();
        r2q[i].assign(rl, 0);                                    // std::map::operator[] reads an absent key as 0
        for (const auto &kv : r.refToQuery)
            if (kv.first < rl) r2q[i][kv.first] = (int32_t)kv.second;
        std::memset(&extra[i], 0, sizeof(dnb_read_extra));
        extra[i].ref_to_query = r2q[i].data();
        extra[i].is_reverse = r.isReverse ? 1 : 0;
        extra[i].ref_start = (uint32_t)r.refStart;
        extra[i].ref_end = (uint32_t)r.refEnd;
    }
    dnb_batch *b = nullptr;
    int rc = dnb_submit_llr(ctx, descs.data(), extra.data(), n, w, &b);
    if (rc != DNB_OK) die("dnb_submit_llr", rc);
    note_device(b);
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        dnb_read_result o;
        if ((rc = dnb_result(b, i, &o)) != DNB_OK) die("dnb_result", rc);
        unpack_result(r, o);
        if (r.eventAlignment.empty()) continue;                  // detect.cpp:879-883: failed read, no llAcrossRead
        dnb_analogue_result a;
        if ((rc = dnb_batch_analogue_result(b, i, &a)) != DNB_OK) die("dnb_batch_analogue_result", rc);
        r.humanReadable_detectOut = ">" + r.readID + " " + r.referenceMappedTo + " " + std::to_string(r.refStart) + " " +
                                    std::to_string(r.refEnd) + " " + (r.isReverse ? "rev" : "fwd") + "\n";
        for (uint32_t s = 0; s < a.n_sites; s++) {
            if (a.n_events[s] == 0) continue;                    // no call for this site (:442, :515)
            const unsigned int p = a.pos_on_ref[s], q = r.refToQuery.at(p);
            std::string kmerQuery = r.basecall.substr(q - k / 2, k), kmerRef = r.referenceSeqMappedTo.substr(p - k / 2, k);
            int globalPos = r.refStart + (int)p;
            if (r.isReverse) {
                globalPos = r.refEnd - (int)p - 1;
                kmerQuery = reverseComplement(kmerQuery);
                kmerRef = reverseComplement(kmerRef);
            }
            const double llr = a.log_analogue[s] - a.log_thymidine[s];   // detect.cpp:548
            r.humanReadable_detectOut += std::to_string(globalPos) + "\t" + std::to_string(llr) + "\t" + kmerRef + "\t" + kmerQuery + "\n";
            r.refCoordToCalls[globalPos] = std::make_pair(llr, 0.);
        }
    }
    dnb_release(b);
}

}  // namespace dnb_shim

void llAcrossRead(DNAscent::read &r, unsigned int windowLength) {
    std::vector<DNAscent::read *> one(1, &r);
    dnb_shim::llAcrossRead_batch(one, windowLength);
}

double sequenceProbability(std::vector<double> &observations, std::string &sequence, size_t windowSize, bool useBrdU,
                           PoreParameters scalings, size_t BrdUStart, size_t BrdUEnd) {
    const size_t k = Pore_Substrate_Config.kmer_len;
    // the device kernel scores the analogue span llAcrossRead uses (detect.cpp:544-545), its only caller in the
    // reference; any other span is refused loudly rather than scored differently (documented in include/dnascent_b200.h)
    if (useBrdU && (BrdUStart != windowSize - k / 2 || BrdUEnd != windowSize + k / 2))
        throw std::invalid_argument("dnascent_b200 shim: sequenceProbability supports BrdUStart/End = window -/+ k/2 only");
    if (sequence.size() != 2 * windowSize + k)
        throw std::invalid_argument("dnascent_b200 shim: sequence must hold 2*windowSize + k bases");
    uint64_t off[2] = {0, observations.size()};
    double la = 0., lt = 0.;
    int rc = dnb_sequence_probability_batch(dnb_shim::context(), observations.data(), off, sequence.data(), &scalings.shift,
                                            &scalings.scale, &scalings.eventsPerBase, 1, (uint32_t)windowSize, &la, &lt);
    if (rc != DNB_OK) die("dnb_sequence_probability_batch", rc);
    return useBrdU ? la : lt;
}
// ---- eventalign (src/alignment.h:22, src/alignment.cpp:547-744) ------------------------------------------------------
// The window chain and every builtinViterbi run on the device (dnb_eventalign_batch); what stays here is what the
// reference does with the resulting state labels: the text of humanReadable_eventalignOut (alignment.cpp:553, 676-736)
// and r.addSignal (alignment.cpp:723), per raw sample of each recorded event.
namespace dnb_shim {

void eventalign_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength) {
    const unsigned int k = Pore_Substrate_Config.kmer_len;
    const size_t n = reads.size();
    if (n == 0) return;
    std::vector<dnb_eventalign_desc> descs(n);
    std::vector<std::vector<int32_t>> r2q(n);
    std::vector<std::vector<uint32_t>> pairs(n);
    std::vector<std::vector<float>> evm(n);
    std::vector<uint64_t> rec_off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        const size_t rl = r.referenceSeqMappedTo.size();
        r2q[i].assign(rl, 0);                                    // std::map::operator[] reads an absent key as 0
        for (const auto &kv : r.refToQuery)
            if (kv.first < rl) r2q[i][kv.first] = (int32_t)kv.second;
        pairs[i].resize(2 * r.eventAlignment.size());
        for (size_t j = 0; j < r.eventAlignment.size(); j++) {
            pairs[i][2 * j] = r.eventAlignment[j].first;
            pairs[i][2 * j + 1] = r.eventAlignment[j].second;
        }
        evm[i].resize(r.events.size());
        for (size_t j = 0; j < r.events.size(); j++) evm[i][j] = (float)r.events[j].mean;   // float32-exact (event_detection.c:226)
        dnb_eventalign_desc &d = descs[i];
        d.ref = r.referenceSeqMappedTo.data();
        d.ref_len = (uint32_t)rl;
        d.ref_to_query = r2q[i].data();
        d.align_pairs = pairs[i].data();
        d.n_align = (uint32_t)r.eventAlignment.size();
        d.event_mean = evm[i].data();
        d.n_events = (uint32_t)r.events.size();
        d.shift = r.scalings.shift;
        d.scale = r.scalings.scale;
        d.events_per_base = r.scalings.eventsPerBase;
        rec_off[i + 1] = rec_off[i] + r.eventAlignment.size() + 64;
    }
    std::vector<dnb_eventalign_rec> recs(rec_off[n]);
    std::vector<uint32_t> n_rec(n);
    std::vector<int> status(n);
    int rc = dnb_eventalign_batch(context(), descs.data(), n, totalWindowLength, recs.data(), rec_off.data(), n_rec.data(),
                                  status.data());
    if (rc != DNB_OK) die("dnb_eventalign_batch", rc);
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        if (status[i] == DNB_READ_UNDEFINED) throw NegativeLog();          // what eln() does in the reference (alignment.cpp:208)
        if (status[i] != DNB_READ_OK) die("dnb_eventalign_batch (per-read capacity)", DNB_ERR_NOMEM);
        std::string &out = r.humanReadable_eventalignOut;
        out = ">" + r.readID + " " + r.referenceMappedTo + " " + std::to_string(r.refStart) + " " + std::to_string(r.refEnd) + " " + r.strand + "\n";
        const dnb_eventalign_rec *rr = recs.data() + rec_off[i];
        for (uint32_t q = 0; q < n_rec[i]; q++) {
            const unsigned int ref_pos = rr[q].ref_pos;
            std::string kmerStrand = r.referenceSeqMappedTo.substr(ref_pos, k);   // kmer2index takes a non-const reference
            unsigned int event_coord;
            std::string kmerRef;
            if (r.isReverse) {                                             // alignment.cpp:646-648, 690-697
                event_coord = r.refEnd - ref_pos - k / 2 - 1;
                kmerRef = reverseComplement(kmerStrand);
            } else {
                event_coord = r.refStart + ref_pos + k / 2;
                kmerRef = kmerStrand;
            }
            const unsigned int event_indexRef = ref_pos + k / 2;
            const unsigned int event_indexQuery = r.refToQuery.at(event_indexRef);
            const std::vector<double> &raw = r.events[rr[q].event].raw;
            if (rr[q].label == DNB_EA_MATCH) {
                const std::pair<double, double> meanStd = Pore_Substrate_Config.pore_model[kmer2index(kmerStrand, k)];
                const bool called = r.refCoordToCalls.count(event_coord) > 0;
                for (size_t idx_raw = 0; idx_raw < raw.size(); idx_raw++) {
                    const double scaledEvent = (raw[idx_raw] - r.scalings.shift) / r.scalings.scale;
                    out += std::to_string(event_coord) + "\t" + kmerRef + "\t" + std::to_string(scaledEvent) + "\t" + kmerStrand + "\t" +
                           std::to_string(meanStd.first);
                    if (called) {
                        out += "\t" + std::to_string(r.refCoordToCalls.at(event_coord).first) + "\t" +
                               std::to_string(r.refCoordToCalls.at(event_coord).second) + "\n";
                    } else {
                        out += "\n";
                        r.addSignal(kmerStrand, event_coord, event_indexQuery, event_indexRef, scaledEvent, rr[q].indel_score);
                    }
                }
            } else {
                for (size_t idx_raw = 0; idx_raw < raw.size(); idx_raw++) {
                    const double scaledEvent = (raw[idx_raw] - r.scalings.shift) / r.scalings.scale;
                    out += std::to_string(event_coord) + "\t" + kmerRef + "\t" + std::to_string(scaledEvent) + "\t" + std::string(k, 'N') + "\t" + "0" + "\n";
                }
            }
        }
        r.QCpassed = true;                                                 // alignment.cpp:743
    }
}

// eventalign + the DNN input tensors in one device pass (SURVEY s.8 row f2)
void eventalign_features_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength,
                               std::vector<DnnInputs> &out) {
    const size_t n = reads.size();
    out.assign(n, DnnInputs());
    if (n == 0) return;
    std::vector<dnb_eventalign_desc> descs(n);
    std::vector<dnb_feature_desc> feats(n);
    std::vector<std::vector<int32_t>> r2q(n);
    std::vector<std::vector<uint32_t>> pairs(n), starts(n), called(n);
    std::vector<std::vector<float>> evm(n), raw(n);
    std::vector<uint64_t> rec_off(n + 1, 0), pos_off(n + 1, 0);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        const size_t rl = r.referenceSeqMappedTo.size();
        r2q[i].assign(rl, 0);                                    // std::map::operator[] reads an absent key as 0
        for (const auto &kv : r.refToQuery)
            if (kv.first < rl) r2q[i][kv.first] = (int32_t)kv.second;
        pairs[i].resize(2 * r.eventAlignment.size());
        for (size_t j = 0; j < r.eventAlignment.size(); j++) {
            pairs[i][2 * j] = r.eventAlignment[j].first;
            pairs[i][2 * j + 1] = r.eventAlignment[j].second;
        }
        // the samples each event owns, in event order (these are consecutive slices of r.raw, event_handling.cpp:549-575;
        // taking them from r.events keeps this independent of that layout); float32-exact, so narrowing is lossless
        evm[i].resize(r.events.size());
        starts[i].resize(r.events.size() + 1);
        uint32_t at = 0;
        for (size_t j = 0; j < r.events.size(); j++) {
            evm[i][j] = (float)r.events[j].mean;
            starts[i][j] = at;
            at += (uint32_t)r.events[j].raw.size();
        }
        starts[i][r.events.size()] = at;
        raw[i].resize(at);
        for (size_t j = 0, t = 0; j < r.events.size(); j++)
            for (double v : r.events[j].raw) raw[i][t++] = (float)v;
        called[i].reserve(r.refCoordToCalls.size());
        for (const auto &kv : r.refCoordToCalls) called[i].push_back(kv.first);   // std::map: ascending
        dnb_eventalign_desc &d = descs[i];
        d.ref = r.referenceSeqMappedTo.data();
        d.ref_len = (uint32_t)rl;
        d.ref_to_query = r2q[i].data();
        d.align_pairs = pairs[i].data();
        d.n_align = (uint32_t)r.eventAlignment.size();
        d.event_mean = evm[i].data();
        d.n_events = (uint32_t)r.events.size();
        d.shift = r.scalings.shift;
        d.scale = r.scalings.scale;
        d.events_per_base = r.scalings.eventsPerBase;
        dnb_feature_desc &f = feats[i];
        f.raw_pA = raw[i].data();
        f.raw_dac = nullptr;
        f.dac_offset = 0.f; f.dac_scale = 1.f;
        f.n_samples = at;
        f.event_start = starts[i].data();
        f.is_reverse = r.isReverse ? 1 : 0;
        f.ref_start = (uint32_t)r.refStart;
        f.ref_end = (uint32_t)r.refEnd;
        f.called = called[i].empty() ? nullptr : called[i].data();
        f.n_called = (uint32_t)called[i].size();
    }
    for (size_t i = 0; i < n; i++) {
        rec_off[i + 1] = rec_off[i] + reads[i]->eventAlignment.size() + 64;
        const size_t rl = reads[i]->referenceSeqMappedTo.size();
        pos_off[i + 1] = pos_off[i] + (rl >= 9 ? rl - 8 : 0) + 1;
    }
    const size_t rows = pos_off[n];
    std::vector<float> signal(rows * DNB_RAWDEPTH), core(rows), residual(rows);
    std::vector<uint32_t> coords(rows), ri(rows), qi(rows), n_rec(n), n_pos(n);
    std::vector<int32_t> qual(rows);
    std::vector<int> status(n);
    dnb_feature_tensors t = {signal.data(), core.data(), residual.data(), coords.data(), ri.data(), qi.data(), qual.data()};
    int rc = dnb_eventalign_features_batch(context(), descs.data(), feats.data(), n, totalWindowLength, nullptr, rec_off.data(),
                                           n_rec.data(), status.data(), &t, pos_off.data(), n_pos.data());
    if (rc != DNB_OK) die("dnb_eventalign_features_batch", rc);
    for (size_t i = 0; i < n; i++) {
        if (status[i] == DNB_READ_UNDEFINED) throw NegativeLog();          // what eln() does in the reference (alignment.cpp:208)
        if (status[i] != DNB_READ_OK) die("dnb_eventalign_features_batch (per-read capacity)", DNB_ERR_NOMEM);
        DnnInputs &o = out[i];
        const size_t lo = pos_off[i], P = n_pos[i];
        o.signal.assign(signal.begin() + lo * DNB_RAWDEPTH, signal.begin() + (lo + P) * DNB_RAWDEPTH);
        o.core.assign(core.begin() + lo, core.begin() + lo + P);
        o.residual.assign(residual.begin() + lo, residual.begin() + lo + P);
        o.refCoords.assign(coords.begin() + lo, coords.begin() + lo + P);
        o.refIndices.assign(ri.begin() + lo, ri.begin() + lo + P);
        o.queryIndices.assign(qi.begin() + lo, qi.begin() + lo + P);
        o.alignmentQuality.assign(qual.begin() + lo, qual.begin() + lo + P);
        o.QCpassed = true;
        reads[i]->QCpassed = true;                                          // alignment.cpp:743
    }
}

// normaliseEvents + eventalign + tensors on one resident batch
void normalise_eventalign_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength,
                                std::vector<DnnInputs> &out) {
    const size_t n = reads.size();
    out.assign(n, DnnInputs());
    if (n == 0) return;
    dnb_ctx *ctx = context();
    std::vector<Staged> staged(n);
    std::vector<dnb_read_desc> descs(n);
    std::vector<dnb_read_extra> extra(n);
    std::vector<std::vector<int32_t>> r2q(n);
    std::vector<std::vector<uint32_t>> called(n);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        stage_read(r, staged[i], descs[i]);
        const size_t rl = r.referenceSeqMappedTo.size();
        r2q[i].assign(rl, 0);                                    // std::map::operator[] reads an absent key as 0
        for (const auto &kv : r.refToQuery)
            if (kv.first < rl) r2q[i][kv.first] = (int32_t)kv.second;
        for (const auto &kv : r.refCoordToCalls) called[i].push_back(kv.first);
        dnb_read_extra &x = extra[i];
        x.ref_to_query = r2q[i].data();
        x.is_reverse = r.isReverse ? 1 : 0;
        x.ref_start = (uint32_t)r.refStart;
        x.ref_end = (uint32_t)r.refEnd;
        x.called = called[i].empty() ? nullptr : called[i].data();
        x.n_called = (uint32_t)called[i].size();
    }
    dnb_batch *b = nullptr;
    int rc = dnb_submit_chain(ctx, descs.data(), extra.data(), n, totalWindowLength, 0, &b);   // thread-safe, staged like dnb_submit
    if (rc != DNB_OK) die("dnb_submit_chain", rc);
    note_device(b);
    bool negative_log = false;
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < n; i++) {
        DNAscent::read &r = *reads[i];
        dnb_read_result o;
        int rc2 = dnb_result(b, i, &o);
        if (rc2 != DNB_OK) die("dnb_result", rc2);
        unpack_result(r, o);
        if (r.eventAlignment.empty()) continue;                              // detect.cpp:879-881: failed read, no eventalign
        dnb_feature_result f;
        if ((rc2 = dnb_batch_feature_result(b, i, &f)) != DNB_OK) die("dnb_batch_feature_result", rc2);
        if (f.status == DNB_READ_UNDEFINED) { negative_log = true; continue; }
        if (f.status != DNB_READ_OK) die("dnb_batch_eventalign_features (per-read capacity)", DNB_ERR_NOMEM);
        DnnInputs &d = out[i];
        d.signal.assign(f.signal, f.signal + (size_t)f.n_pos * DNB_RAWDEPTH);
        d.core.assign(f.core, f.core + f.n_pos);
        d.residual.assign(f.residual, f.residual + f.n_pos);
        d.refCoords.assign(f.coords, f.coords + f.n_pos);
        d.refIndices.assign(f.ref_index, f.ref_index + f.n_pos);
        d.queryIndices.assign(f.query_index, f.query_index + f.n_pos);
        d.alignmentQuality.assign(f.quality, f.quality + f.n_pos);
        d.QCpassed = true;
        r.QCpassed = true;                                                   // alignment.cpp:743
    }
    dnb_release(b);
    if (negative_log) throw NegativeLog();                                   // what eln() does in the reference (alignment.cpp:208)
}

}  // namespace dnb_shim

void eventalign(DNAscent::read &r, unsigned int totalWindowLength) {
    std::vector<DNAscent::read *> one{&r};
    dnb_shim::eventalign_batch(one, totalWindowLength);
}
#endif  // DNB_SHIM_WITH_HMM
