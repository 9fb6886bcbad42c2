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
