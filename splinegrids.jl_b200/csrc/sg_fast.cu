// sg_fast.cu -- tiled fast paths (to be filled in; currently everything routes to the generic kernels)
#include "sg_fast.cuh"

template <typename T>
int sg_evaluate_fast(T *, const SgGridArgs<T> &, const T *, const T *, cudaStream_t) { return SG_ERR_UNSUPPORTED; }

template <typename T>
int sg_evaluate_adjoint_fast(T *, const SgGridArgs<T> &, const SgSpanStarts<T> &, SgAdjointHeader *, const T *,
                             const T *, void *, cudaStream_t) { return SG_ERR_UNSUPPORTED; }

size_t sg_adjoint_fast_scratch_bytes(int, const int64_t *, const int64_t *, int, const int *, int) { return 0; }

template int sg_evaluate_fast<float>(float *, const SgGridArgs<float> &, const float *, const float *, cudaStream_t);
template int sg_evaluate_fast<double>(double *, const SgGridArgs<double> &, const double *, const double *, cudaStream_t);
template int sg_evaluate_adjoint_fast<float>(float *, const SgGridArgs<float> &, const SgSpanStarts<float> &, SgAdjointHeader *,
                                             const float *, const float *, void *, cudaStream_t);
template int sg_evaluate_adjoint_fast<double>(double *, const SgGridArgs<double> &, const SgSpanStarts<double> &, SgAdjointHeader *,
                                              const double *, const double *, void *, cudaStream_t);
