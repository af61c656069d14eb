// Declarations for the operator-resident CG kernel (cg_resident.cu).
#pragma once
#include "krylov_cta.cuh"

namespace lxb {
bool cg_resident_applicable(const KrylovParams<float>& p);
int cg_resident_launch(KrylovParams<float> p, cudaStream_t st);
}  // namespace lxb
