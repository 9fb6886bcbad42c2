"""Small-case driver for compute-sanitizer (memcheck / racecheck / synccheck): one sweep incl. TAU_M / Tau_p of Hubbard 4x4,
Kondo 2x2 and a projector run with 2 chains each, the blocked QR at 256 and CGR2_2 at 64."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from alf_b200.api import AlfB200
from alf_b200 import api
from alf_b200.model import hubbard_square, kondo_square, z2_matter_square
for model, nw in ((hubbard_square(4, 4, 0.6), 3), (kondo_square(2, 2, 0.4), 2), (hubbard_square(4, 4, 0.4, projector=True, theta=0.2, trial="dimer"), 2)):
    g = AlfB200(model, n_chains=2, nwrap=nw); g.set_seeds([5, 6]); g.fields_set(); g.init_sweep(); g.sweep(1, 1); print(model.name, g.control()["XMAXG"]); g.close()
rng = np.random.default_rng(0)
A = rng.normal(size=(1, 256, 256)); api.test_qdrp_blocked(A, False); print("qr256 ok")
U = np.linalg.qr(rng.normal(size=(1, 64, 64)))[0]; api.test_cgr2_2(U, np.ones((1, 64)), U, U, np.ones((1, 64)), U, 0, False); print("cgr22 ok")
# round-1 additions: Ising action tables + star moves (staged G, one-barrier PlaceGR steps), continuous fields, UDV_Wrap_Pivot
for model, nw in ((z2_matter_square(4, 4, 0.3, projector=True, theta=0.2, g=0.8, K=0.5, J=0.7, h=0.9), 2), (z2_matter_square(4, 4, 0.3, propose_s0=True), 2),
                  (hubbard_square(4, 4, 0.4, continuous=True), 2), (hubbard_square(4, 4, 0.4, Mz=False, continuous=True), 2)):
    g = AlfB200(model, n_chains=2, nwrap=nw); g.set_seeds([5, 6]); g.fields_set(); g.init_sweep(); g.sweep(1, 0); print(model.name, g.control()["XMAXG"]); g.close()
os.environ["ALF_B200_NO_STAGE_G"] = "1"
model = z2_matter_square(4, 4, 0.3); g = AlfB200(model, n_chains=2, nwrap=2); g.set_seeds([5, 6]); g.fields_set(); g.init_sweep(); g.sweep(1, 0); print("z2 unstaged", g.control()["XMAXG"]); g.close()
del os.environ["ALF_B200_NO_STAGE_G"]
A = rng.normal(size=(2, 24, 9)) + 1j * rng.normal(size=(2, 24, 9)); api.udv_wrap_pivot(A, True); api.udv_wrap_pivot(rng.normal(size=(1, 256, 256)), False); print("udv_wrap_pivot ok")
# Compute_Fermion_Det, Langevin forces / update, HMC on continuous fields (Mz real and SU(2) complex)
for mz in (True, False):
    model = hubbard_square(4, 4, 0.4, Mz=mz, continuous=True)
    g = AlfB200(model, n_chains=2, nwrap=2); g.set_seeds([5, 6]); g.fields_set(); g.init_sweep()
    ld, ph = g.compute_fermion_det(); g.langevin_update(0.02, 1.5); acc, w = g.hmc_update(0.05, 2); print("det / langevin / hmc", ld[0], w); g.close()
# round-2 additions: ring-group op kernel (persistent, double-buffered cp.async panels; rings of 16 = FULL path, rings of 4 = predicated path), register-panel QR,
# block-diagonal observable kernel, scalar tables, time-dependent couplings, 8 x 8 (rings of 8)
from alf_b200.model import obs_scal_tables
for (l1, l2, beta, nw) in ((16, 16, 0.3, 2), (8, 8, 0.3, 2), (8, 4, 0.3, 2)):
    model = hubbard_square(l1, l2, beta)
    g = AlfB200(model, n_chains=2, nwrap=nw); g.set_seeds([5, 6]); g.fields_set(); g.set_obs_scal_tables(obs_scal_tables(model)); g.init_sweep()
    g.obs_eq_enable(True); g.obs_tau_enable(True); g.sweep(1, 1); print(model.name, l1, l2, g.control()["XMAXG"], g.obs()[:6]); g.close()
model = hubbard_square(4, 4, 0.4)
for n, row in enumerate(model.Op_V):
    for op in row:
        op.g_t = np.array([op.g * (1.0 + 0.2 * np.cos(nt + n)) for nt in range(model.Ltrot)], dtype=np.complex128)
g = AlfB200(model, n_chains=2, nwrap=2); g.set_seeds([5, 6]); g.fields_set(); g.init_sweep(); g.sweep(1, 1); print("g_t", g.control()["XMAXG"]); g.close()
