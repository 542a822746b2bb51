// Kernel instantiations: multi-exponential correlator functors, K = 1..4 and the cumsum(dE) variant.
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_MULTIEXP, "multiexp", MultiExp<1>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<2>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<3>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<4>),
    B200LM_ENTRY(F_MULTIEXP_DE, "multiexp_de", MultiExpDE<1>),
    B200LM_ENTRY(F_MULTIEXP_DE, "multiexp_de", MultiExpDE<2>),
    B200LM_ENTRY_TEAM(F_MULTIEXP_DE, "multiexp_de", MultiExpDE<3>),
    B200LM_ENTRY_TEAM(F_MULTIEXP_DE, "multiexp_de", MultiExpDE<4>),
};
const FunctorEntry* registry_multiexp(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
