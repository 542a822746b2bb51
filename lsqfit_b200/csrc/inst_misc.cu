// Kernel instantiations: the remaining example models.
#define B200LM_DEFINE_ENTRIES
#include "registry.h"
namespace b200lm {
typedef ADFunctor<SimpleBody, 2, 2> Simple;
typedef OffsetExpModel OffsetExp;
static const FunctorEntry kEntries[] = {
    B200LM_ENTRY(F_SIMPLE, "simple", Simple),
    B200LM_ENTRY(F_OFFSET_EXP, "offset_exp", OffsetExp),
    B200LM_ENTRY(F_POLY, "poly", Poly<1>),
    B200LM_ENTRY(F_POLY, "poly", Poly<2>),
    B200LM_ENTRY(F_POLY, "poly", Poly<3>),
    B200LM_ENTRY(F_POLY, "poly", Poly<4>),
    B200LM_ENTRY(F_POLY, "poly", Poly<5>),
    B200LM_ENTRY(F_POLY, "poly", Poly<6>),
    B200LM_ENTRY(F_EXP_POLY, "exp_poly", ExpPoly<3>),
    B200LM_ENTRY(F_EXP_POLY, "exp_poly", ExpPoly<4>),
    B200LM_ENTRY(F_XERR_LOGISTIC, "xerr_logistic", XerrLogistic<15>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<1>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<2>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<3>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<4>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<5>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<6>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<7>),
    B200LM_ENTRY(F_GATHER, "gather", Gather<8>),
};
const FunctorEntry* registry_misc(int* n) { *n = sizeof(kEntries) / sizeof(kEntries[0]); return kEntries; }
}  // namespace b200lm
