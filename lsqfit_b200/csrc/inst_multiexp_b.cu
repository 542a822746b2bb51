// Kernel instantiations: multi-exponential correlator functors, K = 5..8.
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<5>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<6>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<7>),
    B200LM_ENTRY_TEAM(F_MULTIEXP, "multiexp", MultiExp<8>),
};
const FunctorEntry* registry_multiexp_b(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
