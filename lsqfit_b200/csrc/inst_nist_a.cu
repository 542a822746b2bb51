// Kernel instantiations: NIST StRD functors (first half).
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
typedef ADFunctor<Misra1aBody, 2> Misra1a;
typedef ADFunctor<ChwirutBody, 3> Chwirut;
typedef ADFunctor<LanczosBody, 6> Lanczos;
typedef ADFunctor<GaussBody, 8> Gauss;
typedef ADFunctor<DanwoodBody, 2> Danwood;
typedef ADFunctor<Misra1bBody, 2> Misra1b;
typedef ADFunctor<Misra1cBody, 2> Misra1c;
typedef ADFunctor<Misra1dBody, 2> Misra1d;
typedef ADFunctor<Kirby2Body, 5> Kirby2;
typedef ADFunctor<Hahn1Body, 7> Hahn1;
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_MISRA1A, "misra1a", Misra1a),
    B200LM_ENTRY(F_CHWIRUT, "chwirut", Chwirut),
    B200LM_ENTRY(F_LANCZOS, "lanczos", Lanczos),
    B200LM_ENTRY(F_GAUSS, "gauss", Gauss),
    B200LM_ENTRY(F_DANWOOD, "danwood", Danwood),
    B200LM_ENTRY(F_MISRA1B, "misra1b", Misra1b),
    B200LM_ENTRY(F_MISRA1C, "misra1c", Misra1c),
    B200LM_ENTRY(F_MISRA1D, "misra1d", Misra1d),
    B200LM_ENTRY(F_KIRBY2, "kirby2", Kirby2),
    B200LM_ENTRY(F_HAHN1, "hahn1", Hahn1),
};
const FunctorEntry* registry_nist_a(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
