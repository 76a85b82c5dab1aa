// tests/cpp/dropin/spmm_opt.h — the reference-side patch of INTEGRATION.md §2, verbatim: put ahead of
// PA4/handout/include on the include path, it replaces the handout's spmm_opt.h so that the reference's
// UNMODIFIED test/test_spmm.cu (`new SpMMOpt(g, kLen)`) instantiates the B200 engine.
#ifndef SpMM_OPT_H
#define SpMM_OPT_H
#define SPMM_B200_WITH_HANDOUT   // derive from the handout's own class SpMM (spmm_base.h)
#include "spmm_b200.hpp"
typedef SpMMB200 SpMMOpt;
#endif
