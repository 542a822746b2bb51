// Kernel instantiations: larger polynomials (tests/test_lsqfit.py:878-880 fits 25 coefficients).
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_POLY, "poly", Poly<8>),
    B200LM_ENTRY(F_POLY, "poly", Poly<10>),
    B200LM_ENTRY(F_POLY, "poly", Poly<12>),
    B200LM_ENTRY(F_POLY, "poly", Poly<16>),
    B200LM_ENTRY(F_POLY, "poly", Poly<20>),
    B200LM_ENTRY(F_POLY, "poly", Poly<25>),
};
const FunctorEntry* registry_poly_b(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
