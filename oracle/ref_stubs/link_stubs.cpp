// TEST SCAFFOLDING for oracle/_ref only.
// Link-time stand-ins for symbols that the reference translation units mention but that the
// oracle never calls: TensorFlow C API, pod5/fast5 signal readers, the TF model loaders and the
// handful of htslib entry points used by BAM I/O (the oracle hand-builds bam1_t records, so
// htslib itself is not built).  Calling any of them is a bug -> abort().
#include <cstdio>
#include <cstdlib>
#include "detect.h"
#include "pod5.h"
#include "fast5.h"

#define DNB_UNREACHABLE(name) do { std::fprintf(stderr, "oracle/_ref: stub %s called\n", name); std::abort(); } while (0)

extern "C" {
TF_Status *TF_NewStatus(void) { return nullptr; }
void TF_DeleteStatus(TF_Status *) {}
TF_Code TF_GetCode(const TF_Status *) { return TF_OK; }
const char *TF_Message(const TF_Status *) { return ""; }
TF_Graph *TF_NewGraph(void) { DNB_UNREACHABLE("TF_NewGraph"); }
void TF_DeleteGraph(TF_Graph *) {}
TF_Operation *TF_GraphOperationByName(TF_Graph *, const char *) { DNB_UNREACHABLE("TF_GraphOperationByName"); }
TF_SessionOptions *TF_NewSessionOptions(void) { DNB_UNREACHABLE("TF_NewSessionOptions"); }
void TF_DeleteSessionOptions(TF_SessionOptions *) {}
void TF_SetConfig(TF_SessionOptions *, const void *, size_t, TF_Status *) {}
TF_Session *TF_LoadSessionFromSavedModel(const TF_SessionOptions *, const TF_Buffer *, const char *,
                                         const char *const *, int, TF_Graph *, TF_Buffer *, TF_Status *) {
    DNB_UNREACHABLE("TF_LoadSessionFromSavedModel");
}
void TF_DeleteSession(TF_Session *, TF_Status *) {}
void TF_DeleteBuffer(TF_Buffer *) {}
void TF_DeleteImportGraphDefOptions(TF_ImportGraphDefOptions *) {}
TF_Tensor *TF_NewTensor(TF_DataType, const int64_t *, int, void *, size_t, void (*)(void *, size_t, void *), void *) {
    DNB_UNREACHABLE("TF_NewTensor");
}
void TF_DeleteTensor(TF_Tensor *) {}
TF_DataType TF_TensorType(const TF_Tensor *) { return TF_FLOAT; }
size_t TF_TensorByteSize(const TF_Tensor *) { return 0; }
void *TF_TensorData(const TF_Tensor *) { return nullptr; }
void TF_SessionRun(TF_Session *, const TF_Buffer *, const TF_Output *, TF_Tensor *const *, int, const TF_Output *,
                   TF_Tensor **, int, const TF_Operation *const *, int, TF_Buffer *, TF_Status *) {
    DNB_UNREACHABLE("TF_SessionRun");
}
}  // extern "C"

void pod5_getSignal(DNAscent::read &) { DNB_UNREACHABLE("pod5_getSignal"); }
void fast5_getSignal(DNAscent::read &) { DNB_UNREACHABLE("fast5_getSignal"); }
std::pair<std::shared_ptr<ModelSession>, std::shared_ptr<TF_Graph *>> model_load_cpu_twoInputs(const char *, unsigned int) {
    DNB_UNREACHABLE("model_load_cpu_twoInputs");
}
std::pair<std::shared_ptr<ModelSession>, std::shared_ptr<TF_Graph *>> model_load_gpu_twoInputs(const char *, unsigned char,
                                                                                                unsigned int) {
    DNB_UNREACHABLE("model_load_gpu_twoInputs");
}
