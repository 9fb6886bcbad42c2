// T = cplx instantiation of the engine (SU(N) Hubbard, Kondo, ...)
#define ALF_T cplx
#define ALF_NAME(x) x##_cplx
#include "alf_inst.inc"
