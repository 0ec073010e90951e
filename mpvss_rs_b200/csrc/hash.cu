// __global__ entry points of the device-side transcript hashing (sha2_dev.cuh): one SHA-256 chain per thread
// over the framed row of one share, and the single-thread whole-box chain kept as a measured alternative.
#include "hash_launch.h"

namespace shadev {

constexpr int TPB = 64;  // 64 rounds fully unrolled with a 16-word window: register-heavy, short-lived threads

__global__ void __launch_bounds__(TPB) row_hash_kernel(RowHashArgs A) { row_hash_body(A, blockIdx.x * TPB + threadIdx.x); }
__global__ void __launch_bounds__(32) box_hash_kernel(BoxHashArgs A) { box_hash_body(A, threadIdx.x); }

cudaError_t launch_row_hash(const RowHashArgs& A, cudaStream_t s) {
  if (A.n == 0 || A.slot_stride <= 8) return cudaErrorInvalidValue;
  row_hash_kernel<<<(A.n + TPB - 1) / TPB, TPB, 0, s>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_box_hash(const BoxHashArgs& A, cudaStream_t s) {
  if (A.n == 0 || A.slot_stride <= 8) return cudaErrorInvalidValue;
  box_hash_kernel<<<1, 32, 0, s>>>(A);
  return cudaGetLastError();
}

}  // namespace shadev
