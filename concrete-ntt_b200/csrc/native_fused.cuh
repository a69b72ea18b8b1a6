// native_fused.cuh -- fused negacyclic polymul (placeholder until the fused kernel lands below).
#pragma once
namespace cntt {
cudaError_t native_polymul_fused(const NativePlanDev&, void*, const void*, const void*, size_t, cudaStream_t)
{
    return cudaErrorNotSupported;
}
} // namespace cntt
