"""Bin output of scalar observables in ALF's text layout, so the unchanged Analysis tools read it.

Print_bin_Vec (Prog/observables_mod.F90:772-870): the accumulators are divided by the number of measurements N, summed
over ranks and divided by the number of ranks (:788-806, here: chains), then ONE record is appended to `<name>_scal`:
    I10 (N_obs + 1),  N_obs x ' (' E25.17E3 ',' E25.17E3 ')',  E26.17E3 (average sign)            (:864-868)
and `<name>_scal_info` is created once with the analysis mode (:847-861).
"""
from __future__ import annotations

import os

import numpy as np


def _e(x: float, width: int, digits: int = 17, expw: int = 3) -> str:
    """Fortran Ew.dEe edit descriptor: [-]0.ddd...dE+eee with the mantissa in [0.1, 1), right-justified in `width`."""
    if not np.isfinite(x):
        return ("NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")).rjust(width)
    if x == 0.0:
        body, ex = "0" * digits, 0
    else:
        m, e = f"{abs(x):.{digits - 1}e}".split("e")      # d.ddd...d (digits significant digits), correctly rounded
        body, ex = m.replace(".", ""), int(e) + 1
    out = ("-" if (x < 0 or (x == 0 and np.signbit(x))) else "") + "0." + body + "E" + ("+" if ex >= 0 else "-") + f"{abs(ex):0{expw}d}"
    return out.rjust(width)


def format_scal_record(obs_vec, ave_sign: float) -> str:
    obs_vec = np.atleast_1d(np.asarray(obs_vec, dtype=np.complex128))
    line = f"{obs_vec.size + 1:10d}"
    for z in obs_vec:
        line += " (" + _e(z.real, 25) + "," + _e(z.imag, 25) + ")"
    return line + _e(float(ave_sign), 26)


def print_bin_vec(file_pr: str, obs_sum, sign_sum: float, n_meas_per_chain: float, n_chains: int,
                  analysis_mode: str = "identity", description=None) -> str:
    """obs_sum / sign_sum: accumulators summed over all chains (after the NCCL reduction); n_meas_per_chain = Obs%N."""
    obs_vec = np.asarray(obs_sum, dtype=np.complex128) / float(n_meas_per_chain) / float(n_chains)
    ave_sign = float(sign_sum) / float(n_meas_per_chain) / float(n_chains)
    info = file_pr + "_info"
    if not os.path.exists(info):
        with open(info, "w") as f:
            f.write("====== Analysis Mode ======\n" + analysis_mode + "\n")
            if description:
                f.write("====== Description ======\n" + "\n".join(description) + "\n")
    rec = format_scal_record(obs_vec, ave_sign)
    with open(file_pr, "a") as f:
        f.write(rec + "\n")
    return rec


def read_scal(file_pr: str):
    """What Analysis/ana_mod.F90:526-560 (read_vec) does with a `_scal` file: returns (obs[bin, n], sign[bin])."""
    obs, sign = [], []
    for ln in open(file_pr):
        ln = ln.rstrip("\n")
        if not ln.strip():
            continue
        n = int(ln[:10]) - 1
        rest = ln[10:]
        vals = []
        for k in range(n):
            a = rest.index("("); b = rest.index(")")
            re_s, im_s = rest[a + 1:b].split(",")
            vals.append(complex(float(re_s), float(im_s)))
            rest = rest[b + 1:]
        obs.append(vals); sign.append(float(rest))
    return np.asarray(obs), np.asarray(sign)


# ----------------------------------------------------------------------------------------------- lattice observables
def fourier_r_to_k(x_r, latt):
    """FT_R_to_K_C (Libraries/Modules/lattices_v3_mod.F90:847-876): X(k) = 1/N sum_r exp(-i k.r) X(r) on the k-points
    k = m1 b1_p + m2 b2_p, b_p = 2 pi / L (square-type Bravais lattice with orthogonal unit vectors: listk enumerates like list)."""
    pts = np.asarray(latt.list, dtype=np.float64)                      # (N, 2) integer coordinates of r and of k
    kv = pts * np.array([2.0 * np.pi / latt.L1, 2.0 * np.pi / latt.L2])
    phase = np.exp(-1j * (kv @ pts.T))                                 # [k, r]
    return (phase @ np.asarray(x_r, dtype=np.complex128)) / latt.N, kv


def _write_info(file_pr, channel, ntau, dtau, latt, norb, n_coord, orb_pos):
    """`<file>_info`, written once (Prog/observables_mod.F90:457-491 and :664-698, same text for both record types)."""
    info = file_pr + "_info"
    if orb_pos is None:
        orb_pos = np.zeros((norb, 2))
    orb_pos = np.asarray(orb_pos, dtype=np.float64).reshape(norb, -1)
    if not os.path.exists(info):
        with open(info, "w") as f:           # formats 11-15 of observables_mod.F90:462-466
            f.write(f"{'Observable':>20s}: {os.path.basename(file_pr)}\n{'Channel':>20s}: {channel}\n{'Ntau':>20s}: {ntau:10d}\n")
            f.write(f"{'dtau':>20s}: " + _e(dtau or 0.0, 26) + "\n       ====== Bravais Lattice ======\n")
            f.write(f"{'Unit cells':>20s}: {latt.N:10d}\n{'L1':>20s}: " + _e(float(latt.L1), 26) + _e(0.0, 26) + "\n")
            f.write(f"{'L2':>20s}: " + _e(0.0, 26) + _e(float(latt.L2), 26) + "\n")
            f.write(f"{'a1':>20s}: " + _e(1.0, 26) + _e(0.0, 26) + f"\n{'a2':>20s}: " + _e(0.0, 26) + _e(1.0, 26) + "\n")
            f.write("       ========= Unit cell =========\n")
            f.write(f"{'Coordination number':>20s}: {int(n_coord):10d}\n{'Number of orbitals':>20s}: {norb:10d}\n{'Ndim':>20s}: {orb_pos.shape[1]:10d}\n")
            for no in range(norb):
                f.write(f"{'Orbital %d' % (no + 1):>20s}: " + "".join(_e(x, 26) for x in orb_pos[no]) + "\n")


def print_bin_latt(name, obs_sum, bg_sum, sign_sum, n_meas_per_chain, n_chains, latt, dtau=None, channel="---", n_coord=2, orb_pos=None):
    """Print_bin_Latt (Prog/observables_mod.F90:355-515), text layout (:494-512): appends one bin to `<name>_tau` (or `<name>_eq` if
    there is a single time point).  obs_sum[nt, no, no1, r]: real-space accumulator summed over chains (Obs_Latt(imj, nt, no, no1));
    bg_sum[no]: Obs_Latt0 summed over chains and time points; n_meas_per_chain = Obs%N.  n_coord, orb_pos[no][:] = Latt_unit%N_coord and
    Latt_unit%Orb_pos_p (Prog/Predefined_Latt_mod.F90:117-237; default: square lattice, one orbital at the origin): they go into the
    `_info` file (:457-491) in exactly the order Analysis/ana_mod.F90:285-309 reads them back."""
    obs_sum = np.asarray(obs_sum, dtype=np.complex128); ntau, norb, _, ns = obs_sum.shape
    assert ns == latt.N
    suffix = "_eq" if ntau == 1 else "_tau"
    file_pr = name + suffix
    norm = float(n_meas_per_chain) * float(n_chains)
    obs = obs_sum / norm                                               # Obs_Latt / N, averaged over ranks
    bg = np.asarray(bg_sum, dtype=np.complex128) / (norm * ns * ntau)  # Obs_Latt0 / (N Ns Ntau)
    ave_sign = float(sign_sum) / norm
    _write_info(file_pr, channel, ntau, dtau, latt, norb, n_coord, orb_pos)
    lines = []
    if ntau == 1:
        lines.append(_e(ave_sign, 25) + f"{norb:11d}{latt.N:11d}")
    else:
        lines.append(_e(ave_sign, 25) + f"{norb:11d}{latt.N:11d}{ntau:11d}" + _e(dtau or 0.0, 26))
    for no in range(norb):
        lines.append("(" + _e(bg[no].real, 25) + "," + _e(bg[no].imag, 25) + ")")
    xk = None
    obs_k = np.zeros_like(obs)
    for nt in range(ntau):
        for no in range(norb):
            for no1 in range(norb):
                obs_k[nt, no, no1], xk = fourier_r_to_k(obs[nt, no, no1], latt)
    for i in range(latt.N):
        lines.append(_e(xk[i, 0], 25) + " " + _e(xk[i, 1], 25))
        for nt in range(ntau):
            for no in range(norb):
                for no1 in range(norb):
                    z = obs_k[nt, no, no1, i]
                    lines.append("(" + _e(z.real, 25) + "," + _e(z.imag, 25) + ")")
    with open(file_pr, "a") as f:
        f.write("\n".join(lines) + "\n")
    return file_pr


def print_bin_latt_local(name, obs_sum, sign_sum, n_meas_per_chain, n_chains, latt, dtau=None, channel="---", n_coord=2, orb_pos=None):
    """Print_bin_Latt_Local (Prog/observables_mod.F90:578-767), text layout (:700-717): appends one bin of a site-resolved
    observable to `<name>_local` (one time point) or `<name>_localtau`.  obs_sum[nt, no, r] = Obs_Latt(I, nt, no) summed over
    chains; no Fourier transform and no background: each unit cell is headed by its real-space position list(I,1) a1 + list(I,2) a2."""
    obs_sum = np.asarray(obs_sum, dtype=np.complex128); ntau, norb, ns = obs_sum.shape
    assert ns == latt.N
    file_pr = name + ("_local" if ntau == 1 else "_localtau")
    norm = float(n_meas_per_chain) * float(n_chains)
    obs = obs_sum / norm
    ave_sign = float(sign_sum) / norm
    _write_info(file_pr, channel, ntau, dtau, latt, norb, n_coord, orb_pos)
    if ntau == 1:
        lines = [_e(ave_sign, 25) + f"{norb:11d}{latt.N:11d}"]
    else:
        lines = [_e(ave_sign, 25) + f"{norb:11d}{latt.N:11d}{ntau:11d}" + _e(dtau or 0.0, 26)]
    pts = np.asarray(latt.list, dtype=np.float64)
    for i in range(latt.N):
        lines.append(_e(pts[i, 0], 25) + " " + _e(pts[i, 1], 25))
        for nt in range(ntau):
            for no in range(norb):
                z = obs[nt, no, i]
                lines.append("(" + _e(z.real, 25) + "," + _e(z.imag, 25) + ")")
    with open(file_pr, "a") as f:
        f.write("\n".join(lines) + "\n")
    return file_pr


def read_latt_local(file_pr):
    """Reads `_local` / `_localtau` bins back the way Analysis/ana_mod.F90 (read_latt with no background and one orbital index) does:
    returns (sign[nb], obs[nb, ntau, norb, r], x_r[r, 2])."""
    def cplx(s):
        a, b = s.strip()[1:-1].split(",")
        return complex(float(a), float(b))
    with open(file_pr) as f:
        rows = [ln.rstrip("\n") for ln in f if ln.strip()]
    signs, bins, pos = [], [], None
    p = 0
    while p < len(rows):
        head = rows[p].split(); p += 1
        norb, ns = int(head[1]), int(head[2]); ntau = int(head[3]) if len(head) > 3 else 1
        signs.append(float(head[0]))
        obs = np.zeros((ntau, norb, ns), dtype=np.complex128); xr = np.zeros((ns, 2))
        for i in range(ns):
            xr[i] = [float(t) for t in rows[p].split()]; p += 1
            for nt in range(ntau):
                for no in range(norb):
                    obs[nt, no, i] = cplx(rows[p]); p += 1
        bins.append(obs); pos = xr
    return np.array(signs), np.array(bins), pos


def read_latt(file_pr):
    """What Analysis/ana_mod.F90:57-220 (read_latt) does with a `_tau` / `_eq` file: returns a list of bins
    (sign, bg[norb], k[N, 2], obs[k, nt, no, no1])."""
    toks = open(file_pr).read().split("\n")
    pos = 0; bins = []

    def cplx(s):
        a, b = s.strip()[1:-1].split(","); return complex(float(a), float(b))
    while pos < len(toks) and toks[pos].strip():
        hdr = toks[pos].split(); pos += 1
        sign, norb, n = float(hdr[0]), int(hdr[1]), int(hdr[2]); ntau = int(hdr[3]) if len(hdr) > 3 else 1
        bg = [cplx(toks[pos + k]) for k in range(norb)]; pos += norb
        ks = np.zeros((n, 2)); obs = np.zeros((n, ntau, norb, norb), dtype=np.complex128)
        for i in range(n):
            ks[i] = [float(x) for x in toks[pos].split()]; pos += 1
            for nt in range(ntau):
                for no in range(norb):
                    for no1 in range(norb):
                        obs[i, nt, no, no1] = cplx(toks[pos]); pos += 1
        bins.append((sign, np.array(bg), ks, obs))
    return bins


def read_latt_info(file_pr):
    """The read sequence of Analysis/ana_mod.F90:285-309 on `<file>_info`, line by line with its formats 11-13 (A22 label, then the value):
    returns (channel, ntau, dtau, n_unit, L1_p, L2_p, a1_p, a2_p, n_coord, norb, orb_pos[norb, ndim_unit])."""
    with open(file_pr + "_info") as f:
        ln = f.read().split("\n")
    it = iter(ln)

    def a22(line):
        return line[22:]

    def reals(line):
        body = a22(line); return [float(body[i:i + 26]) for i in range(0, len(body.rstrip()), 26)]
    next(it)                                             # read(10, *)
    channel = a22(next(it)).strip()
    ntau = int(a22(next(it))[:10])
    dtau = reals(next(it))[0]
    next(it)
    n_unit = int(a22(next(it))[:10])
    L1_p = reals(next(it)); L2_p = reals(next(it)); a1_p = reals(next(it)); a2_p = reals(next(it))
    next(it)
    n_coord = int(a22(next(it))[:10]); norb = int(a22(next(it))[:10]); ndim_unit = int(a22(next(it))[:10])
    orb = np.zeros((norb, ndim_unit))
    for no in range(norb):
        v = reals(next(it)); assert len(v) == ndim_unit
        orb[no] = v
    return channel, ntau, dtau, n_unit, L1_p, L2_p, a1_p, a2_p, n_coord, norb, orb
