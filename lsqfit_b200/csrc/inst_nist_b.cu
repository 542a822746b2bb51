// Kernel instantiations: NIST StRD functors (second half).
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
typedef ADFunctor<NelsonBody, 3, 2> Nelson;
typedef ADFunctor<Mgh17Body, 5> Mgh17;
typedef ADFunctor<Roszman1Body, 4> Roszman1;
typedef ADFunctor<EnsoBody, 9> Enso;
typedef ADFunctor<Mgh09Body, 4> Mgh09;
typedef ADFunctor<Rat42Body, 3> Rat42;
typedef ADFunctor<Mgh10Body, 3> Mgh10;
typedef ADFunctor<Eckerle4Body, 3> Eckerle4;
typedef ADFunctor<Rat43Body, 4> Rat43;
typedef ADFunctor<Bennett5Body, 3> Bennett5;
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_NELSON, "nelson", Nelson),
    B200LM_ENTRY(F_MGH17, "mgh17", Mgh17),
    B200LM_ENTRY(F_ROSZMAN1, "roszman1", Roszman1),
    B200LM_ENTRY(F_ENSO, "enso", Enso),
    B200LM_ENTRY(F_MGH09, "mgh09", Mgh09),
    B200LM_ENTRY(F_RAT42, "rat42", Rat42),
    B200LM_ENTRY(F_MGH10, "mgh10", Mgh10),
    B200LM_ENTRY(F_ECKERLE4, "eckerle4", Eckerle4),
    B200LM_ENTRY(F_RAT43, "rat43", Rat43),
    B200LM_ENTRY(F_BENNETT5, "bennett5", Bennett5),
};
const FunctorEntry* registry_nist_b(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
