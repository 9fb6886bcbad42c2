// Host driver in C++ that plays the role of Prog/main.F90 (lines 589-906: storage fill, G(0), bins of sweeps, measurements, precision / acceptance
// report) through the C-ABI of include/alf_b200.h ONLY -- no Python, no torch, no oracle.  It is what the Fortran shim does on the Fortran side
// ("test the C side with a C++ host driver that plays the role of main.F90").
//
// usage: main_driver <model.bin> <n_chains> <nwrap> <n_bins> <n_sweeps> <ltau> <out.bin>
// model.bin (written by tests/test_gpu_cabi_driver.py from alf_b200.model): the PUBLIC fields of Hamiltonian_main / type Operator as plain arrays --
// exactly what alf_b200_shim.F90 flattens out of Op_V, Op_T: header int32 {ndim, n_fl, n_sun, ltrot, n_opv, n_opt, symm}, then per Op_V(n, nf) and
// Op_T(nc, nf): int32 {N, n_non_zero, diag, type}, int32 P[N], complex128 U[N*N], float64 E[N], complex128 g, alpha; then int32 seeds[n_chains].
// out.bin: per bin the 16 scalar observables, then the final control vector (16), the phase of every chain (complex) and G(:, :, nf = 1) of chain 0.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <complex>
#include "../../include/alf_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != ALF_OK) { std::fprintf(stderr, "%s failed: %d (%s)\n", #x, rc_, h ? alf_b200_last_error(h) : ""); return 2; } } while (0)

template <typename X> static bool rd(FILE* f, X* p, size_t n) { return std::fread(p, sizeof(X), n, f) == n; }

int main(int argc, char** argv) {
  if (argc != 8) { std::fprintf(stderr, "usage: %s model.bin n_chains nwrap n_bins n_sweeps ltau out.bin\n", argv[0]); return 1; }
  const int n_chains = std::atoi(argv[2]), nwrap = std::atoi(argv[3]), n_bins = std::atoi(argv[4]), n_sweeps = std::atoi(argv[5]), ltau = std::atoi(argv[6]);
  FILE* f = std::fopen(argv[1], "rb"); if (!f) { std::perror("model.bin"); return 1; }
  int32_t hd[7]; if (!rd(f, hd, 7)) return 1;
  const int ndim = hd[0], n_fl = hd[1], n_sun = hd[2], ltrot = hd[3], n_opv = hd[4], n_opt = hd[5], symm = hd[6];
  alf_b200_handle* h = nullptr;
  CHECK(alf_b200_create(&h, ndim, n_fl, n_sun, ltrot, nwrap, n_opv, n_opt, symm, 0, n_chains, 0));
  for (int pass = 0; pass < 2; ++pass) {                       // Op_V(n, nf), then Op_T(nc, nf)
    const int cnt = pass == 0 ? n_opv : n_opt;
    for (int n = 1; n <= cnt; ++n) for (int nf = 1; nf <= n_fl; ++nf) {
      int32_t q[4]; if (!rd(f, q, 4)) return 1;
      std::vector<int32_t> P(q[0]); std::vector<std::complex<double>> U((size_t)q[0] * q[0]); std::vector<double> E(q[0]); std::complex<double> ga[2];
      if (!rd(f, P.data(), P.size()) || !rd(f, U.data(), U.size()) || !rd(f, E.data(), E.size()) || !rd(f, ga, 2)) return 1;
      if (pass == 0) CHECK(alf_b200_set_op_v(h, n, nf, q[0], q[1], q[2], q[3], P.data(), reinterpret_cast<double*>(U.data()), E.data(), ga[0].real(), ga[0].imag(), ga[1].real(), ga[1].imag()));
      else CHECK(alf_b200_set_op_t(h, n, nf, q[0], q[2], P.data(), reinterpret_cast<double*>(U.data()), E.data(), ga[0].real(), ga[0].imag()));
    }
  }
  std::vector<int32_t> seeds(n_chains); if (!rd(f, seeds.data(), seeds.size())) return 1;
  std::fclose(f);
  CHECK(alf_b200_finalize_model(h));                           // Hop_mod_init + Op_set tables
  CHECK(alf_b200_set_seeds(h, seeds.data()));                  // Set_Random_number_Generator, one stream per chain (= per MPI rank)
  CHECK(alf_b200_fields_set(h));                               // nsigma%in: random start (Fields_set)
  CHECK(alf_b200_init_sweep(h));                               // main.F90:589-631
  FILE* out = std::fopen(argv[7], "wb"); if (!out) { std::perror("out.bin"); return 1; }
  std::vector<double> obs(alf_b200_obs_size(h));
  for (int nb = 0; nb < n_bins; ++nb) {                        // DO NBC = 1, NBIN (main.F90:642)
    CHECK(alf_b200_obs_reset(h));                              // ham%Init_obs
    CHECK(alf_b200_sweep(h, n_sweeps, ltau));                  // DO NSW = 1, NSWEEP: the sequential sweep of all chains (main.F90:714-887)
    CHECK(alf_b200_reduce_bins(h, 0));                         // MPI_REDUCE of Print_bin_* (a no-op on one rank)
    CHECK(alf_b200_get_obs(h, obs.data()));                    // ham%Pr_obs
    std::fwrite(obs.data(), sizeof(double), obs.size(), out);
  }
  double ctl[16]; CHECK(alf_b200_reduce_control(h, 0, ctl));   // Control_Print
  std::fwrite(ctl, sizeof(double), 16, out);
  std::vector<double> ph(2 * (size_t)n_chains), G(2 * (size_t)ndim * ndim);
  CHECK(alf_b200_get_phase(h, ph.data())); CHECK(alf_b200_get_green(h, 0, 1, 0, G.data()));
  std::fwrite(ph.data(), sizeof(double), ph.size(), out); std::fwrite(G.data(), sizeof(double), G.size(), out);
  std::fclose(out);
  std::printf("acceptance %.6f  precision Green mean %.3e max %.3e  sweeps %d x %d chains\n", ctl[8] / (ctl[7] > 0 ? ctl[7] : 1), ctl[0] / (ctl[2] > 0 ? ctl[2] : 1), ctl[1], n_bins * n_sweeps, n_chains);
  CHECK(alf_b200_destroy(h));
  return 0;
}
