// T = double instantiation of the engine (real models: Mz-Hubbard, ...)
#define ALF_T double
#define ALF_NAME(x) x##_real
#include "alf_inst.inc"
