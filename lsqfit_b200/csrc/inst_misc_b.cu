// Kernel instantiations: the spline model of examples/spline.py and the
// shared-energy composite models of simultaneous (MultiFitter) correlator fits.
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_SPLINE_POLY, "spline_poly", SplinePoly<4, 5>),      // examples/spline.py
    B200LM_ENTRY(F_SPLINE_POLY, "spline_poly", SplinePoly<4, 0>),
    B200LM_ENTRY(F_SPLINE_POLY, "spline_poly", SplinePoly<6, 0>),
    B200LM_ENTRY(F_MULTIEXP_SHARED2, "multiexp_shared2", MultiExpShared<1, 2>),
    B200LM_ENTRY(F_MULTIEXP_SHARED2, "multiexp_shared2", MultiExpShared<2, 2>),
    B200LM_ENTRY(F_MULTIEXP_SHARED2, "multiexp_shared2", MultiExpShared<3, 2>),
    B200LM_ENTRY(F_MULTIEXP_SHARED2, "multiexp_shared2", MultiExpShared<4, 2>),
    B200LM_ENTRY(F_MULTIEXP_SHARED3, "multiexp_shared3", MultiExpShared<2, 3>),
    B200LM_ENTRY(F_MULTIEXP_SHARED3, "multiexp_shared3", MultiExpShared<3, 3>),
};
const FunctorEntry* registry_misc_b(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
