/* TEST SCAFFOLDING for oracle/_ref only (never shipped, never on the product path).
 * Declarations-only stand-in for the slice of the TensorFlow C API that
 * /root/reference/src/{tensor.h,detect.cpp} mention, so those translation units can be
 * compiled for their NON-TensorFlow functions (sequenceProbability, llAcrossRead).
 * None of these functions is ever called by the oracle harness. */
#ifndef DNB_ORACLE_TF_STUB_H
#define DNB_ORACLE_TF_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct TF_Status TF_Status;
typedef struct TF_Graph TF_Graph;
typedef struct TF_Tensor TF_Tensor;
typedef struct TF_Session TF_Session;
typedef struct TF_SessionOptions TF_SessionOptions;
typedef struct TF_Operation TF_Operation;
typedef struct TF_ImportGraphDefOptions TF_ImportGraphDefOptions;
typedef struct TF_Buffer {
    const void *data;
    size_t length;
    void (*data_deallocator)(void *data, size_t length);
} TF_Buffer;
typedef struct TF_Output {
    TF_Operation *oper;
    int index;
} TF_Output;
typedef enum TF_Code { TF_OK = 0 } TF_Code;
typedef enum TF_DataType { TF_FLOAT = 1 } TF_DataType;

TF_Status *TF_NewStatus(void);
void TF_DeleteStatus(TF_Status *);
TF_Code TF_GetCode(const TF_Status *);
const char *TF_Message(const TF_Status *);
TF_Graph *TF_NewGraph(void);
void TF_DeleteGraph(TF_Graph *);
TF_Operation *TF_GraphOperationByName(TF_Graph *, const char *);
TF_SessionOptions *TF_NewSessionOptions(void);
void TF_DeleteSessionOptions(TF_SessionOptions *);
void TF_SetConfig(TF_SessionOptions *, const void *, size_t, TF_Status *);
TF_Session *TF_LoadSessionFromSavedModel(const TF_SessionOptions *, const TF_Buffer *, const char *,
                                         const char *const *, int, TF_Graph *, TF_Buffer *, TF_Status *);
void TF_DeleteSession(TF_Session *, TF_Status *);
void TF_DeleteBuffer(TF_Buffer *);
void TF_DeleteImportGraphDefOptions(TF_ImportGraphDefOptions *);
TF_Tensor *TF_NewTensor(TF_DataType, const int64_t *dims, int num_dims, void *data, size_t len,
                        void (*deallocator)(void *, size_t, void *), void *deallocator_arg);
void TF_DeleteTensor(TF_Tensor *);
TF_DataType TF_TensorType(const TF_Tensor *);
size_t TF_TensorByteSize(const TF_Tensor *);
void *TF_TensorData(const TF_Tensor *);
void TF_SessionRun(TF_Session *, const TF_Buffer *run_options, const TF_Output *inputs,
                   TF_Tensor *const *input_values, int ninputs, const TF_Output *outputs,
                   TF_Tensor **output_values, int noutputs, const TF_Operation *const *target_opers,
                   int ntargets, TF_Buffer *run_metadata, TF_Status *);
#ifdef __cplusplus
}
#endif
#endif
