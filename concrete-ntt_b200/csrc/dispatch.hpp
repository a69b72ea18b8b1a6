// dispatch.hpp -- non-template entry points into the per-class kernel instantiations.
// Each arithmetic class is compiled in its own translation unit (inst_*.cu) so the build parallelises.
#pragma once
#include "ntt_kernels.cuh"

namespace cntt {

#define CNTT_DECLARE_CLASS(A)                                                                                      \
    cudaError_t ntt_##A(const PlanDev<A>& pl, typename A::W* data, size_t batch, bool fwd, cudaStream_t st);        \
    cudaError_t ntt_strided_##A(const PlanDev<A>& pl, typename A::W* data, size_t batch, size_t poly_stride, bool fwd,    \
                                cudaStream_t st);                                                                   \
    cudaError_t pointwise_strided_##A(const PlanDev<A>& pl, int op, typename A::W* dst, const typename A::W* a,    \
                                      const typename A::W* b, size_t batch, size_t poly_stride, cudaStream_t st);  \
    cudaError_t pointwise_##A(const PlanDev<A>& pl, int op, typename A::W* dst, const typename A::W* a,            \
                              const typename A::W* b, size_t nwords, cudaStream_t st);                             \
    bool uses_last_##A(int logn);                                                                                  \
    cudaError_t build_last_##A(int logn, bool fwd, const typename A::Tw* heap, typename A::Tw* out, cudaStream_t st);

CNTT_DECLARE_CLASS(A32L4)
CNTT_DECLARE_CLASS(A32L2)
CNTT_DECLARE_CLASS(A32G)
CNTT_DECLARE_CLASS(A64L4)
CNTT_DECLARE_CLASS(A64L2)
CNTT_DECLARE_CLASS(A64S)
CNTT_DECLARE_CLASS(A64G)

#define CNTT_DEFINE_CLASS(A)                                                                                       \
    bool uses_last_##A(int logn) { return plan_uses_last<A>(logn); }                                               \
    cudaError_t build_last_##A(int logn, bool fwd, const typename A::Tw* heap, typename A::Tw* out, cudaStream_t st) \
    {                                                                                                              \
        return launch_build_last<A>(logn, fwd, heap, out, st);                                                          \
    }                                                                                                              \
    cudaError_t ntt_##A(const PlanDev<A>& pl, typename A::W* data, size_t batch, bool fwd, cudaStream_t st)         \
    {                                                                                                              \
        return fwd ? launch_ntt<A, true>(pl, data, batch, st) : launch_ntt<A, false>(pl, data, batch, st);         \
    }                                                                                                              \
    cudaError_t ntt_strided_##A(const PlanDev<A>& pl, typename A::W* data, size_t batch, size_t poly_stride, bool fwd,    \
                                cudaStream_t st)                                                                    \
    {                                                                                                              \
        return fwd ? launch_ntt<A, true>(pl, data, batch, st, poly_stride) : launch_ntt<A, false>(pl, data, batch, st, poly_stride); \
    }                                                                                                              \
    cudaError_t pointwise_strided_##A(const PlanDev<A>& pl, int op, typename A::W* dst, const typename A::W* a,    \
                                      const typename A::W* b, size_t batch, size_t poly_stride, cudaStream_t st)   \
    {                                                                                                              \
        switch (op) {                                                                                              \
        case OP_MUL_ASSIGN_NORMALIZE: return launch_pointwise_strided<A, OP_MUL_ASSIGN_NORMALIZE>(pl, dst, a, b, batch, poly_stride, st); \
        case OP_NORMALIZE: return launch_pointwise_strided<A, OP_NORMALIZE>(pl, dst, a, b, batch, poly_stride, st); \
        case OP_MUL_ACCUMULATE: return launch_pointwise_strided<A, OP_MUL_ACCUMULATE>(pl, dst, a, b, batch, poly_stride, st); \
        default: return cudaErrorInvalidValue;                                                                     \
        }                                                                                                          \
    }                                                                                                              \
    cudaError_t pointwise_##A(const PlanDev<A>& pl, int op, typename A::W* dst, const typename A::W* a,            \
                              const typename A::W* b, size_t nwords, cudaStream_t st)                              \
    {                                                                                                              \
        switch (op) {                                                                                              \
        case OP_MUL_ASSIGN_NORMALIZE: return launch_pointwise<A, OP_MUL_ASSIGN_NORMALIZE>(pl, dst, a, b, nwords, st); \
        case OP_NORMALIZE: return launch_pointwise<A, OP_NORMALIZE>(pl, dst, a, b, nwords, st);                    \
        case OP_MUL_ACCUMULATE: return launch_pointwise<A, OP_MUL_ACCUMULATE>(pl, dst, a, b, nwords, st);          \
        default: return cudaErrorInvalidValue;                                                                     \
        }                                                                                                          \
    }

} // namespace cntt
