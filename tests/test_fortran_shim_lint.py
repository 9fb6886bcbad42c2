"""CPU checks of the Fortran side of the boundary (alf_b200/fortran/): no Fortran compiler exists in the image, so this is what can be verified without
one.  (1) alf_b200_c_api.F90 is regenerated from include/alf_b200.h and must be identical (interfaces cannot drift; every interface body imports every
kind it uses by construction -- the defect of round 1's hand-written block); (2) every C-ABI call in alf_b200_shim.F90 names an existing entry point and
passes a legal number of arguments; (3) the shim exports the reference's procedures with the reference's dummy-argument lists (read
from /root/reference when present, else from the list recorded here); (4) block constructs balance."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_fortran_interface as gen        # noqa: E402

SHIM = os.path.join(ROOT, "alf_b200", "fortran", "alf_b200_shim.F90")
# name -> dummy arguments (lower case), as in the reference (file:line in SURVEY.md 8b)
REFERENCE_SIGNATURES = {
    "wrapur": ["ntau", "ntau1", "udvr"], "wrapul": ["ntau1", "ntau", "udvl"],
    "cgr": ["phase", "nvar", "grup", "udvr", "udvl"], "cgrp": ["phase", "grup", "udvr", "udvl"],
    "cgr2_2": ["grt0", "gr00", "grtt", "gr0t", "udv2", "udv1", "lq"],
    "qdrp_decompose": ["ndim", "n_part", "mat", "d", "ipvt", "tau", "work", "lwork"],
    "udv_wrap_pivot": ["a", "u", "d", "v", "ncon", "n1", "n2"], "decompose_udv_state": ["udvr"],
    "hop_mod_mmthr": ["in", "nf", "t"], "hop_mod_mmthr_m1": ["in", "nf", "t"], "hop_mod_mmthl": ["in", "nf", "t"],
    "hop_mod_mmthl_m1": ["in", "nf", "t"], "hop_mod_mmthlc": ["in", "nf", "t"], "hop_mod_symm": ["out", "in", "t1", "t2"],
    "wrapgrup": ["gr", "ntau", "phase", "propose_s0", "nt_sequential_start", "nt_sequential_end", "n_global_tau"],
    "wrapgrdo": ["gr", "ntau", "phase", "propose_s0", "nt_sequential_start", "nt_sequential_end", "n_global_tau"],
    "tau_m": ["udvst", "gr", "phase", "nstm", "nwrap", "stab_nt", "lobs_st", "lobs_en"],
}
REFERENCE_FILES = {"wrapur": "Prog/wrapur_mod.F90", "wrapul": "Prog/wrapul_mod.F90", "cgr": "Prog/cgr1_mod.F90", "cgrp": "Prog/cgr1_mod.F90", "cgr2_2": "Prog/cgr2_2_mod.F90",
                   "qdrp_decompose": "Prog/QDRP_decompose_mod.F90", "udv_wrap_pivot": "Prog/UDV_WRAP_mod.F90", "decompose_udv_state": "Prog/udv_state_mod.F90",
                   "hop_mod_mmthr": "Prog/Hop_mod.F90", "hop_mod_mmthr_m1": "Prog/Hop_mod.F90", "hop_mod_mmthl": "Prog/Hop_mod.F90", "hop_mod_mmthl_m1": "Prog/Hop_mod.F90",
                   "hop_mod_mmthlc": "Prog/Hop_mod.F90", "hop_mod_symm": "Prog/Hop_mod.F90", "wrapgrup": "Prog/Wrapgr_mod.F90", "wrapgrdo": "Prog/Wrapgr_mod.F90", "tau_m": "Prog/tau_m_mod.F90"}


def joined_source(path):
    """Fortran source with comments stripped and continuation lines joined."""
    out, cur = [], ""
    for ln in open(path):
        if ln.lstrip().startswith("#"):
            continue
        code = ln.split("!")[0].rstrip() if '"' not in ln and "'" not in ln else re.sub(r"\s!\s.*$", "", ln.rstrip())
        if not code.strip():
            continue
        code = code.strip()
        if code.startswith("&"):
            code = code[1:].lstrip()
        if code.endswith("&"):
            cur += code[:-1] + " "; continue
        out.append(cur + code); cur = ""
    return out


def subroutine_args(lines):
    sigs = {}
    for ln in lines:
        m = re.match(r"(?i)\s*(?:recursive\s+)?subroutine\s+(\w+)\s*(?:\((.*?)\))?\s*$", ln)
        if m:
            sigs[m.group(1).lower()] = [a.strip().lower() for a in (m.group(2) or "").split(",") if a.strip()]
    return sigs


def test_interface_module_is_generated_from_the_header():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_interface.py"), "--check"]) == 0
    txt = open(gen.OUT).read()
    assert txt.count("end function") == len(gen.prototypes()) >= 80
    # every interface body imports the kinds it uses (the round-1 defect)
    for body in re.findall(r"(?s)function \w+\(.*?end function", txt):
        imported = set(re.search(r"import :: (.*)", body).group(1).replace(" ", "").split(","))
        used = set(re.findall(r"\b(c_\w+)\b", re.sub(r"import :: .*", "", body)))
        assert used <= imported, (body[:60], used - imported)


def split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([":
            depth += 1
        if ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_shim_calls_match_the_c_prototypes():
    protos = {name: params for _, name, params in gen.prototypes()}
    src = " ".join(joined_source(SHIM))
    n_calls = 0
    for m in re.finditer(r"\b(alf_b200_\w+)\s*\(", src):
        name = m.group(1)
        if name in ("alf_b200_attach", "alf_b200_detach", "alf_b200_batched_sweep", "alf_b200_reduce"):
            continue
        i = m.end(); depth = 1; j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0); j += 1
        args = split_args(src[i:j - 1])
        assert name in protos, f"{name} is not declared in include/alf_b200.h"
        npos = sum(1 for a in args if not re.match(r"^\w+\s*=[^=]", a)); nkw = len(args) - npos
        pointer_params = sum(1 for ct, _ in protos[name] if ct.endswith("*") and "alf_b200_handle" not in ct)
        assert npos + nkw <= len(protos[name]) and len(protos[name]) - (npos + nkw) <= pointer_params, (name, args, len(protos[name]))
        n_calls += 1
    assert n_calls >= 25


def test_shim_exports_the_reference_procedures_with_their_argument_lists():
    sigs = subroutine_args(joined_source(SHIM))
    ref_root = "/root/reference"
    for name, args in REFERENCE_SIGNATURES.items():
        assert name in sigs, f"{name} missing in the shim"
        mine = ["ndim" if a == "ndim_" else a for a in sigs[name]]
        assert mine == args, (name, mine, args)
        path = os.path.join(ref_root, REFERENCE_FILES[name])
        if os.path.exists(path):      # in the build container: the recorded list is itself checked against the reference's source
            ref = subroutine_args(joined_source(path))
            assert ref.get(name) == args, (name, ref.get(name), args)
    public = re.search(r"(?is)public :: WRAPUR.*?TAU_M", open(SHIM).read()).group(0).lower()
    for name in REFERENCE_SIGNATURES:
        assert re.search(r"\b%s\b" % name, public), name


def test_shim_block_constructs_balance():
    lines = [ln.lower() for ln in joined_source(SHIM)]
    count = lambda pat: sum(1 for ln in lines if re.match(pat, ln))
    assert count(r"\s*subroutine\s") == count(r"\s*end\s*subroutine")
    assert count(r"\s*module\s+\w+\s*$") == count(r"\s*end\s*module") == 1
    assert count(r"\s*do\s+\w+\s*=") == count(r"\s*end\s*do")
    assert count(r"\s*if\s*\(.*\)\s*then\s*$") == count(r"\s*end\s*if")
    assert count(r"\s*allocate\s*\(") >= 1 and "implicit none" in " ".join(lines)
