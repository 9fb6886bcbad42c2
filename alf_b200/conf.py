"""Field-configuration checkpoints in ALF's plain-text layout (`confout_<rank>` / `confin_<rank>`), one file per chain.

Fields_write_conf / Fields_read_conf (Prog/Fields_mod.F90:750-774, 631-662), non-HDF5 branch:
    line 1      : the K integers of the random-number state, list-directed (`WRITE(10,*) SEED_VEC`)
    then, for NT = 1..Ltrot, I = 1..N_op (I fastest), one list-directed value per line:
        type 1, 2 : nint(real(f(I,NT)))        (integer)
        type 3    : real(f(I,NT))              (double)
        type 4    : f(I,NT)                    (complex, "(re,im)")
File names follow File_i (Libraries/Modules/files_mod.F90): `confout_<rank>`; a run restarts from `confin_<rank>`
(Fields_in, Prog/Fields_mod.F90:372-407), chain c of a handle playing the role of MPI rank c.

The state vector here is the chain's xoshiro256** state as K = 8 32-bit integers (low word first), the seed length of current
libgfortran; like the stream itself it is a declared convention, not libgfortran's scrambled one.
"""
from __future__ import annotations

import os

import numpy as np

SEED_LEN = 8


def state_to_seed_vec(state4) -> np.ndarray:
    """4 x uint64 -> 8 x int32 (low word first), the packing of Ranset (Libraries/Modules/random_wrap_mod.F90:60-77)."""
    s = np.ascontiguousarray(state4, dtype=np.uint64).reshape(4)
    return s.view(np.uint32).astype(np.uint32).view(np.int32).copy()


def seed_vec_to_state(seed_vec) -> np.ndarray:
    v = np.ascontiguousarray(seed_vec, dtype=np.int64).reshape(SEED_LEN)
    return (v & 0xFFFFFFFF).astype(np.uint32).view(np.uint64).copy()


def _fmt_real(x: float) -> str:
    return "  " + repr(float(x))


def write_conf(filename: str, seed_vec, f, types) -> None:
    """f: complex (Ltrot, N_op) = nsigma%f transposed; types: nsigma%t (N_op)."""
    f = np.asarray(f, dtype=np.complex128); types = np.asarray(types, dtype=np.int64)
    assert f.ndim == 2 and f.shape[1] == types.size
    lines = ["".join(f" {int(v):11d}" for v in np.asarray(seed_vec).reshape(-1))]
    for nt in range(f.shape[0]):
        for i in range(f.shape[1]):
            t, z = int(types[i]), f[nt, i]
            if t in (1, 2):
                lines.append(f"{int(np.rint(z.real)):12d}")
            elif t == 3:
                lines.append(_fmt_real(z.real))
            elif t == 4:
                lines.append(f"           ({repr(float(z.real))},{repr(float(z.imag))})")
            # other types: nothing is written (as in the reference)
    with open(filename, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def read_conf(filename: str, ltrot: int, types):
    """Returns (seed_vec int32[K], f complex (Ltrot, N_op)).  List-directed input: values separated by blanks, commas or newlines."""
    types = np.asarray(types, dtype=np.int64); n_op = types.size
    with open(filename) as fh:
        first = fh.readline()
        seed_vec = np.array([int(t) for t in first.replace(",", " ").split()], dtype=np.int64)
        while seed_vec.size < SEED_LEN:                 # a list-directed record may wrap
            seed_vec = np.concatenate([seed_vec, [int(t) for t in fh.readline().replace(",", " ").split()]])
        f = np.zeros((ltrot, n_op), dtype=np.complex128)
        for nt in range(ltrot):
            for i in range(n_op):
                t = int(types[i])
                if t not in (1, 2, 3, 4):
                    continue
                tok = fh.readline().strip()
                if not tok:
                    raise ValueError(f"{filename}: configuration shorter than Ltrot x N_op = {ltrot} x {n_op}")
                if t in (1, 2):
                    f[nt, i] = float(int(tok))
                elif t == 3:
                    f[nt, i] = float(tok.replace("D", "E").replace("d", "e"))
                else:
                    re_, im_ = tok.strip("() ").replace("D", "E").split(",")
                    f[nt, i] = complex(float(re_), float(im_))
    return _to_i32(seed_vec), f


def _to_i32(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.int64)
    return ((v + 2**31) % 2**32 - 2**31).astype(np.int32)


def conf_name(prefix: str, rank: int) -> str:
    return f"{prefix}_{rank}"


def write_confs(g, directory: str, prefix: str = "confout") -> list:
    """One `confout_<chain>` per chain of the handle `g` (AlfB200): its RNG state and field configuration."""
    os.makedirs(directory, exist_ok=True)
    st = g.rng_state().reshape(g.C, 4); f = g.get_fields(); types = np.array([op[0].type for op in g.m.Op_V], dtype=np.int64)       # nsigma%t(n) = Op_V(n,1)%type (Fields_mod.F90:309-330)
    out = []
    for c in range(g.C):
        p = os.path.join(directory, conf_name(prefix, c)); write_conf(p, state_to_seed_vec(st[c]), f[c], types); out.append(p)
    return out


def read_confs(g, directory: str, prefix: str = "confin") -> None:
    """Restart: loads `confin_<chain>` for every chain of the handle (state of the random-number stream and nsigma%f)."""
    types = np.array([op[0].type for op in g.m.Op_V], dtype=np.int64)       # nsigma%t(n) = Op_V(n,1)%type (Fields_mod.F90:309-330)
    st = np.zeros((g.C, 4), dtype=np.uint64); f = np.zeros((g.C, g.m.Ltrot, g.m.n_opv), dtype=np.complex128)
    for c in range(g.C):
        sv, f[c] = read_conf(os.path.join(directory, conf_name(prefix, c)), g.m.Ltrot, types)
        st[c] = seed_vec_to_state(sv)
    g.set_rng_state(st); g.set_fields(f)
