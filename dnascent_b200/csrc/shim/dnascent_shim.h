// dnascent_shim.h -- C++ drop-in layer: the reference's own symbols for the detect signal hot path, re-implemented
// on top of the C ABI of libdnascent_b200 (include/dnascent_b200.h).
//
// Compiled INSIDE the reference tree (it includes the reference's reads.h / config.h) in place of
//   src/event_handling.cpp     normaliseEvents                         (decl. src/event_handling.h:13)
//   src/scrappie/event_detection.c   detect_events                     (decl. src/scrappie/event_detection.h:35)
//   src/probability.cpp        eexp eln lnSum lnProd lnGreaterThan uniformPDF normalPDF cauchyPDF   (src/probability.h:26-33)
// and, with -DDNB_SHIM_WITH_HMM, of the two hot functions of src/detect.cpp
//   sequenceProbability, llAcrossRead                                  (decl. src/detect.h:119,121)
// and of the per-read stage of src/alignment.cpp
//   eventalign (with builtinViterbi)                                   (decl. src/alignment.h:22)
// so that detect.cpp / alignment.cpp / trainCNN.cpp link unchanged (call sites detect.cpp:876, alignment.cpp:856,
// trainCNN.cpp:319).  The batched entry points below are what the patched read loop of detect.cpp:850-908 calls
// (INTEGRATION.md shows the patch); the one-read signatures are kept for every other caller.
#pragma once
#include <vector>

#include "reads.h"   // DNAscent::read, PoreParameters (reference header)

struct dnb_ctx;

namespace dnb_shim {

// Process-wide context on `device` (default: $DNB_DEVICE or 0).  Created on first use; loads the three tables of
// Pore_Substrate_Config (pore_model, unlabelled_model, analogue_model) into HBM.  Thread-safe.
dnb_ctx *context();
void set_device(int device);   // call before the first use
// One process, several GPUs: every batched call below is dealt to the device with the least work in flight
// (dnb_config.devices; also settable as DNB_DEVICES="0,1,...").  Call before the first use.
void set_devices(const std::vector<int> &devices);
unsigned long batches_on_device(int device);   // submissions dealt to that GPU so far (diagnostics)
void shutdown();               // destroys the context (optional; e.g. before pod5_terminate at detect.cpp:917)

// Batched normaliseEvents(r, false): one GPU submission for the whole buffer of reads (the reference's
// `buffer`, detect.cpp:819).  Leaves each read exactly as the reference would: r.events, r.eventAlignment (empty ==
// failed read, detect.cpp:879), r.scalings, r.alignmentQCs.  Reads must have r.raw filled (pod5_getSignal).
void normaliseEvents_batch(const std::vector<DNAscent::read *> &reads, bool useFitPoreModel);

// Batched llAcrossRead (detect.cpp:393-574): the per-site event gathering stays on the host, every forward pass of
// every site of every read runs in one device launch.
void llAcrossRead_batch(const std::vector<DNAscent::read *> &reads, unsigned int windowLength);

// detect --HMM (detect.cpp:876-885): normaliseEvents + llAcrossRead of a whole buffer as ONE device-resident chain
// (dnb_submit_llr): sites, event windows and forward passes are computed on the GPU from the resident alignment.
// Leaves every read as normaliseEvents_batch + llAcrossRead_batch would (failed reads: empty eventAlignment, no calls).
void normalise_llAcrossRead_batch(const std::vector<DNAscent::read *> &reads, unsigned int windowLength);

// Batched eventalign (alignment.cpp:547-744): all window chains and Viterbi passes of the buffer in one device
// launch; humanReadable_eventalignOut and r.addSignal are produced on the host from the returned state records.
void eventalign_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength);

// What runCNN (detect.cpp:586-690) reads from a DNAscent::read after eventalign, as the vectors it would get from
// r.makeSignalTensor() [P*RAWDEPTH], r.makeCoreSequenceTensor(), r.makeResidualSequenceTensor(),
// r.getReferenceCoords(), r.getReferenceIndices(), r.getQueryIndices() and each position's getAlignmentQuality().
struct DnnInputs {
    std::vector<float> signal, core, residual;
    std::vector<unsigned int> refCoords, refIndices, queryIndices;
    std::vector<int> alignmentQuality;
    bool QCpassed = false;         // r.QCpassed as eventalign leaves it (alignment.cpp:743)
};

// `detect`'s use of eventalign (detect.cpp:888): only the r.addSignal side effect is consumed there, never the text.
// This entry runs the window chains AND builds the tensors on the device (dnb_eventalign_features_batch); the raw
// signal goes up as float32 (r.raw is float32-exact, pod5.cpp:60), nothing per-sample is touched on the host.
// r.refCoordToAP is left empty: the patched runCNN takes `out[i]` instead of calling r.make*Tensor().
void eventalign_features_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength,
                               std::vector<DnnInputs> &out);

// The signal part of detect's read loop body (detect.cpp:876-888) as ONE device-resident chain: normaliseEvents ->
// eventalign -> tensors, with the signal, events, alignment and scalings never leaving HBM in between
// (dnb_submit_chain; several buffers may be in flight from different host threads).  Leaves every read as normaliseEvents_batch does
// (a failed read has an empty eventAlignment and gets out[i].QCpassed == false, cf. the `continue` at detect.cpp:879-881)
// and returns the DNN inputs of the others.
void normalise_eventalign_batch(const std::vector<DNAscent::read *> &reads, unsigned int totalWindowLength,
                                std::vector<DnnInputs> &out);

}  // namespace dnb_shim
