// Functor registry: (family, np) -> kernel launchers.  Each inst_*.cu contributes a
// table; capi.cu concatenates them.
#pragma once
#include "lm_types.h"

namespace b200lm {

struct FunctorEntry {
    int family;
    int np;
    int nx;
    const char* name;
    cudaError_t (*fit)(FitParams, int sm_count, size_t smem_budget, cudaStream_t);
    cudaError_t (*resjac)(FitParams, int sm_count, size_t smem_budget, cudaStream_t);
    size_t (*per_warp_bytes)(int rb);
};

const FunctorEntry* registry_multiexp(int* n);
const FunctorEntry* registry_nist_a(int* n);
const FunctorEntry* registry_nist_b(int* n);
const FunctorEntry* registry_misc(int* n);

}  // namespace b200lm

#ifdef B200LM_DEFINE_ENTRIES
#include "lm_kernel.cuh"
#include "functors.cuh"
namespace b200lm {
template <class F>
size_t per_warp_bytes_of(int rb) { return (size_t)FitLayout<F>::per_warp_doubles(rb) * sizeof(double); }
#define B200LM_ENTRY(family, name, ...) \
    { family, __VA_ARGS__::NP, __VA_ARGS__::NX, name, &launch_fit<__VA_ARGS__>, &launch_resjac<__VA_ARGS__>, \
      &per_warp_bytes_of<__VA_ARGS__> }
}  // namespace b200lm
#endif
