// Host-side launchers for the ModpGroup kernels.
#pragma once
#include <cuda_runtime.h>
#include "modp_kernels.cuh"

namespace modp {
cudaError_t launch_horner(int tpi, const HornerArgs& A, cudaStream_t s);
cudaError_t launch_exp2(int tpi, const Exp2Args& A, cudaStream_t s);
cudaError_t launch_mul(int tpi, const MulArgs& A, cudaStream_t s);
}  // namespace modp
