"""`_scal` bin records in ALF's text layout (Prog/observables_mod.F90:864-868) and the normalisation of Print_bin_Vec (:788-806)."""
import numpy as np

from alf_b200.bins import _e, format_scal_record, print_bin_vec, read_scal


def test_fortran_e_descriptor():
    assert _e(1.0, 25) == " 0.10000000000000000E+001"
    assert _e(-0.5, 25) == "-0.50000000000000000E+000"
    assert _e(0.0, 26) == "  0.00000000000000000E+000"
    assert _e(-2.0 ** -10, 25) == "-0.97656250000000000E-003"
    assert _e(0.99999999999999999999, 25) == " 0.10000000000000000E+001"
    assert float(_e(0.1 + 0.2, 25)) == 0.1 + 0.2          # 17 significant digits round-trip a double
    assert len(_e(123456.789, 25)) == 25 and len(_e(-1e-300, 26)) == 26


def test_scal_record_layout_and_roundtrip(tmp_path):
    rec = format_scal_record([1.5 - 2.0j, 0.25j], 0.875)
    assert rec.startswith("         3 ( 0.15") and rec.count("(") == 2 and len(rec) == 10 + 2 * (2 + 25 + 1 + 25 + 1) + 26
    p = str(tmp_path / "Part_scal")
    # two bins: accumulators summed over 4 chains with 199 measurements per chain
    for k in range(2):
        print_bin_vec(p, [4 * 199 * (16.0 + k) + 0.0j], 4 * 199 * 1.0, 199, 4, description=["Particle number"])
    obs, sign = read_scal(p)
    assert obs.shape == (2, 1) and np.allclose(obs[:, 0], [16.0, 17.0]) and np.allclose(sign, 1.0)
    info = open(p + "_info").read()
    assert "====== Analysis Mode ======" in info and "identity" in info and "Particle number" in info


def test_latt_bin_layout_fourier_and_roundtrip(tmp_path):
    """`_tau` records (Prog/observables_mod.F90:494-512): header, backgrounds, then per k-point Ntau*Norb*Norb values; Fourier_R_to_K
    convention X(k) = 1/N sum_r exp(-i k r) X(r) (lattices_v3_mod.F90:847-876)."""
    from alf_b200.bins import fourier_r_to_k, print_bin_latt, read_latt
    from alf_b200.model import Lattice
    latt = Lattice(4, 2); ntau, norb, nchains, nmeas = 3, 1, 5, 2
    r0 = latt.invlist[(0, 0)] - 1; r1 = latt.invlist[(1, 0)] - 1
    x = np.zeros(latt.N, dtype=complex); x[r0] = 8.0
    xk, kv = fourier_r_to_k(x, latt)
    assert np.allclose(xk, 1.0)                                        # delta function at the origin -> constant 8 / N
    x[:] = 0; x[r1] = 8.0
    xk, kv = fourier_r_to_k(x, latt)
    assert np.allclose(xk, np.exp(-1j * kv[:, 0]))                     # shifted by a_1: phase exp(-i k_x)
    obs = np.zeros((ntau, norb, norb, latt.N), dtype=complex)
    for nt in range(ntau):
        obs[nt, 0, 0, r0] = 8.0 * (nt + 1) * nchains * nmeas           # accumulators are sums over chains and measurements
    p = print_bin_latt(str(tmp_path / "Green"), obs, [2.0 * nchains * nmeas * latt.N * ntau], nchains * nmeas * 1.0, nmeas, nchains, latt, dtau=0.1)
    p = print_bin_latt(str(tmp_path / "Green"), obs, [2.0 * nchains * nmeas * latt.N * ntau], nchains * nmeas * 1.0, nmeas, nchains, latt, dtau=0.1)
    assert p.endswith("Green_tau")
    first = open(p).readline()
    assert len(first.rstrip("\n")) == 25 + 3 * 11 + 26 and first.split()[1:4] == ["1", "8", "3"]
    bins = read_latt(p)
    assert len(bins) == 2
    sign, bg, ks, o = bins[1]
    assert sign == 1.0 and np.allclose(bg, [2.0]) and o.shape == (8, 3, 1, 1)
    assert np.allclose(o[:, :, 0, 0], np.array([1.0, 2.0, 3.0])[None, :])
    assert "Unit cells" in open(p + "_info").read()
    # the `_info` file parses with the exact read sequence and formats of Analysis/ana_mod.F90:285-309
    from alf_b200.bins import read_latt_info
    ch, nt_, dt_, nu, L1p, L2p, a1p, a2p, ncoord, norb_, orb = read_latt_info(p)
    assert (ch, nt_, nu, ncoord, norb_) == ("---", 3, 8, 2, 1) and dt_ == 0.1
    assert L1p == [4.0, 0.0] and L2p == [0.0, 2.0] and a1p == [1.0, 0.0] and a2p == [0.0, 1.0] and orb.shape == (1, 2) and not orb.any()
    # two orbitals with three-component positions (Bilayer_square, Predefined_Latt_mod.F90:155-160)
    obs2 = np.zeros((1, 2, 2, latt.N), dtype=complex)
    p2 = print_bin_latt(str(tmp_path / "Den"), obs2, [0.0, 0.0], 1.0, 1, 1, latt, n_coord=2, orb_pos=[[0, 0, 0], [0, 0, -1.0]])
    r = read_latt_info(p2)
    assert r[8:10] == (2, 2) and r[10].shape == (2, 3) and r[10][1, 2] == -1.0 and r[1] == 1
    lines = open(p2 + "_info").read().split("\n")
    assert lines[11].startswith(" Coordination number: ") and lines[12].startswith("  Number of orbitals: ") and lines[13].startswith("                Ndim: ")
    assert lines[14].startswith("           Orbital 1: ") and len(lines[15]) == 22 + 3 * 26


def test_conf_roundtrip_all_field_types(tmp_path):
    """confout/confin text layout (Prog/Fields_mod.F90:631-662, 750-774): seed vector line, then one value per (I, NT), I fastest; integer for
    types 1/2, real for type 3, complex for type 4.  Round trip is exact, and the file parses with the reference's record structure."""
    from alf_b200 import conf
    rng = np.random.default_rng(3); ltrot, types = 5, np.array([1, 2, 3, 4, 1, 3])
    f = np.zeros((ltrot, types.size), dtype=np.complex128)
    f[:, types == 1] = rng.choice([-1, 1], size=(ltrot, 2)); f[:, types == 2] = rng.choice([-2, -1, 1, 2], size=(ltrot, 1))
    f[:, types == 3] = rng.normal(size=(ltrot, 2)); f[:, types == 4] = (rng.normal(size=(ltrot, 1)) + 1j * rng.normal(size=(ltrot, 1)))
    state = rng.integers(0, 2**63, size=4, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    sv = conf.state_to_seed_vec(state)
    assert sv.dtype == np.int32 and sv.size == conf.SEED_LEN and np.array_equal(conf.seed_vec_to_state(sv), state)
    p = str(tmp_path / conf.conf_name("confout", 3)); assert p.endswith("confout_3")
    conf.write_conf(p, sv, f, types)
    lines = open(p).read().splitlines()
    assert len(lines) == 1 + ltrot * types.size and len(lines[0].split()) == conf.SEED_LEN
    assert int(lines[1]) == int(f[0, 0].real) and lines[4].strip().startswith("(")          # I fastest: line 1 + I of slice 1
    sv2, f2 = conf.read_conf(p, ltrot, types)
    assert np.array_equal(sv2, sv) and np.array_equal(f2, f)


def test_latt_local_bin_layout_and_roundtrip(tmp_path):
    """`_local` / `_localtau` records (Prog/observables_mod.F90:700-717): header, then per unit cell its real-space position
    followed by Ntau*Norb values; no Fourier transform, no background line."""
    from alf_b200.bins import print_bin_latt_local, read_latt_local, read_latt_info
    from alf_b200.model import Lattice
    latt = Lattice(4, 2); nchains, nmeas = 3, 7
    rng = np.random.default_rng(5)
    for ntau, suffix, ncol in ((1, "_local", 3), (4, "_localtau", 5)):
        x = rng.standard_normal((ntau, 2, latt.N)) + 1j * rng.standard_normal((ntau, 2, latt.N))
        p = None
        for k in range(2):
            p = print_bin_latt_local(str(tmp_path / "SpinZ"), x * nchains * nmeas * (k + 1), 0.5 * nchains * nmeas, nmeas, nchains, latt,
                                     dtau=0.1, orb_pos=[[0, 0], [0.5, 0.5]])
        assert p.endswith("SpinZ" + suffix)
        rows = open(p).read().split("\n")
        assert len(rows[0].split()) == ncol and len(rows[0]) == 25 + (ncol - 1 if ntau == 1 else 3) * 11 + (0 if ntau == 1 else 26)
        assert len(rows[1]) == 25 + 1 + 25 and rows[2].startswith("(") and len(rows[2]) == 2 * 25 + 3
        assert len(rows) - 1 == 2 * (1 + latt.N * (1 + ntau * 2))
        sign, obs, xr = read_latt_local(p)
        assert np.allclose(sign, 0.5) and obs.shape == (2, ntau, 2, latt.N)
        assert np.allclose(obs[0], x, rtol=1e-15, atol=0) and np.allclose(obs[1], 2 * x, rtol=1e-15, atol=0)
        assert np.array_equal(xr, np.asarray(latt.list, dtype=float))
        info = read_latt_info(p)
        assert info[1] == ntau and info[9] == 2 and info[10][1, 0] == 0.5
