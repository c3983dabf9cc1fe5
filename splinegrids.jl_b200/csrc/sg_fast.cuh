// sg_fast.cuh -- tiled sm_100a fast paths for evaluate! / evaluate_adjoint! (dispatch entry points).
// Each returns SG_ERR_UNSUPPORTED when the shape has no fast path (the caller then runs the generic kernel).
#pragma once
#include "sg_adjoint_generic.cuh"
#include "sg_common.cuh"

template <typename T>
int sg_evaluate_fast(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st);

// What a plan (sg_adjoint_plan_create) read back from the device after the prep kernel.  planned == false: nothing is
// known on the host, every decision is taken on device and the fallback kernels are launched (idle) behind the pipeline.
struct SgAdjKnown {
    bool planned;
    bool fused_ok;     // the column-block tables of the fused double march exist and fit
    int rows2_max;     // largest number of samples in one knot span of dimension 2
    int sf3, sl3;      // first / last span of dimension 3 that holds samples (0: unknown)
    const void *uni;   // host copy of the weights / span starts of dimension 2 (SgM2Uni<T>, sg_adjoint_march2g.cuh) or nullptr
};
size_t sg_m2_uni_bytes(int elem_size);
// fills *uni (SgM2Uni<T>) from host copies of the selected table slice of dimension 2 (n2, P+1) column-major and its span starts
bool sg_m2_uni_fill(void *uni, int elem_size, const void *table2_host, int64_t n2, int P, const int32_t *start2_host, int64_t c2,
                    const int32_t *start3_host, int64_t c3, int span_first3, int span_last3);

template <typename T>
int sg_evaluate_adjoint_fast(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr,
                             const T *eval, const T *weights, void *scratch, const SgAdjKnown &known, cudaStream_t st);

size_t sg_adjoint_fast_scratch_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                     const int *degree, int elem_size);

// Sizes of the column-block tables of the fused double march (ok == false: the shape has no fused pipeline).
struct SgM2gDims {
    bool ok;
    int icap, rmcap, nb1;
    int bw;        // samples of dimension 1 per column block (128 for the 3-D double march, 128 * V for the fused 2-D march)
    int rfast;     // layout of the weights: 1 = [block][li][r] (r fastest, 2-D: lanes walk r), 0 = [block][r][li] (3-D: lanes walk li)
};
SgM2gDims sg_m2g_dims(int nin, const int64_t *n_samples, const int64_t *n_cp, const int *degree, bool rational, int elem_size);

// evaluate! for several derivative orders in one launch (sg_eval_multi.cuh); SG_ERR_UNSUPPORTED: the caller loops
template <typename T>
struct SgMultiArgs;
template <typename T>
int sg_evaluate_multi_fast(const SgMultiArgs<T> &m, int n_der, const SgGridArgs<T> &a, const T *cp, cudaStream_t st);
