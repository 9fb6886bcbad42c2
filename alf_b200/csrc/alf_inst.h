// Entry points of the two template instantiations (alf_inst_real.cu: T = double, alf_inst_cplx.cu: T = cplx).
#pragma once
struct alf_b200_handle; struct EngineBase;
#define ALF_INST_DECL(SUF) \
  EngineBase* alf_make_engine_##SUF(alf_b200_handle* h); \
  void alf_t_qdrp_##SUF(int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases); \
  void alf_t_udv_##SUF(int n, int batch, char side, double* U, double* D, double* V); \
  void alf_t_cgr_##SUF(int n, int batch, int nvar, int stab, const double* UR, const double* DR, const double* VR, const double* UL, const double* DL, \
                       const double* VL, const double* detUR, const double* detUL, double* G, double* phase); \
  void alf_t_cgr22_##SUF(int n, int batch, int stab, const double* U2, const double* D2, const double* V2, const double* U1, const double* D1, \
                         const double* V1, double* out4); \
  void alf_t_qdrp_blk_##SUF(int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases, double* Q); \
  void alf_t_cgrp_##SUF(int n, int n_part, int batch, const double* UR, const double* UL, double* G, double* phase); \
  void alf_t_udv_wrap_pivot_##SUF(int n1, int n2, int batch, const double* A, double* U, double* D, double* V); \
  void alf_t_gemm_##SUF(int ta, int tb, int m, int n, int k, int batch, const double* A, const double* B, double* Cc);
ALF_INST_DECL(real)
ALF_INST_DECL(cplx)
