// TMA-staged warp-per-system LU for 32x32 fp32 systems (lu_tma.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lxb {
bool lu32_tma_eligible(const float* A, int64_t sA, const float* lu, int64_t batch, int n);
// solve: x = A^-1 b fused with the factorisation (lu / piv optional); !solve: factor only.
int lu32_tma_launch(const float* A, int64_t sA, const float* b, int64_t sb, float* x, float* lu,
                    int32_t* piv, int64_t batch, bool solve, cudaStream_t st);
}  // namespace lxb
