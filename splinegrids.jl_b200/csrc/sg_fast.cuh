// sg_fast.cuh -- tiled sm_100a fast paths for evaluate! / evaluate_adjoint! (dispatch entry points).
// Each returns SG_ERR_UNSUPPORTED when the shape has no fast path (the caller then runs the generic kernel).
#pragma once
#include "sg_adjoint_generic.cuh"
#include "sg_common.cuh"

template <typename T>
int sg_evaluate_fast(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st);

template <typename T>
int sg_evaluate_adjoint_fast(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr,
                             const T *eval, const T *weights, void *scratch, cudaStream_t st);

size_t sg_adjoint_fast_scratch_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                     const int *degree, int elem_size);
