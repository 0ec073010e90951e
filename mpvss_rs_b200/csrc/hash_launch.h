// Host-side launchers for the device-side transcript hashing (bodies in sha2_dev.cuh, entry points in hash.cu).
#pragma once
#include <cuda_runtime.h>
#include "sha2_dev.cuh"

namespace shadev {
cudaError_t launch_row_hash(const RowHashArgs& A, cudaStream_t s);
cudaError_t launch_box_hash(const BoxHashArgs& A, cudaStream_t s);
}  // namespace shadev
