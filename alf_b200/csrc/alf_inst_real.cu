// T = double instantiation of the engine (real models: Mz-Hubbard, ...)
#define ALF_T double
#define ALF_NAME(x) x##_real
#include "alf_inst.inc"
#ifdef ALF_QR_PROF
extern "C" void alf_b200_qr_prof_read(unsigned long long* out, int reset) {      // experimental builds only (build.py: ALF_QR_PROF=1)
  cudaMemcpyFromSymbol(out, g_qr_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_qr_prof, z, sizeof(z)); }
}
#endif
#ifdef ALF_UPD_PROF
extern "C" void alf_b200_upd_prof_read(unsigned long long* out, int reset) {      // experimental builds only (build.py: ALF_UPD_PROF=1)
  cudaMemcpyFromSymbol(out, g_upd_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_upd_prof, z, sizeof(z)); }
}
#endif
