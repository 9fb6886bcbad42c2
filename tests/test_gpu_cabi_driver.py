"""GPU test (-m gpu): a compiled C++ host that plays Prog/main.F90 through the C-ABI alone (tests/cabi/main_driver.cpp: model tables, seeds, random
start, storage fill, bins of sweeps, bin reduction, control report) gives the oracle's observables, phases and Green function -- the drop-in
boundary works for a host that is neither Python nor linked against torch."""
import os
import struct
import subprocess

import numpy as np
import pytest

from alf_b200 import build as alf_build
from alf_b200.model import hubbard_square, kondo_square, flatten_ops
from oracle.oracle import Oracle
from common import relF, SEEDS, TOL_G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_model(path, m, seeds):
    ov, ot = flatten_ops(m)
    with open(path, "wb") as f:
        f.write(struct.pack("<7i", m.Ndim, m.N_FL, m.N_SUN, m.Ltrot, m.n_opv, m.n_opt, int(m.Symm)))
        for lst, key in ((ov, "n"), (ot, "nc")):
            for o in sorted(lst, key=lambda o: (o[key], o["nf"])):
                f.write(struct.pack("<4i", o["N"], o.get("nnz", o["N"]), int(o["diag"]), int(o.get("type", 0))))
                f.write(np.asarray(o["P"], dtype=np.int32).tobytes()); f.write(np.asfortranarray(o["U"], dtype=np.complex128).tobytes(order="F"))
                f.write(np.asarray(o["E"], dtype=np.float64).tobytes())
                f.write(np.array([o["g"], o.get("alpha", 0.0)], dtype=np.complex128).tobytes())
        f.write(np.asarray(seeds, dtype=np.int32).tobytes())


@pytest.mark.parametrize("which", ["hubbard_mz", "kondo_complex"])
def test_cpp_host_driver_plays_main_through_the_cabi(tmp_path, which):
    lib = alf_build.build()
    exe = str(tmp_path / "main_driver")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cabi", "main_driver.cpp"), "-o", exe, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"])
    m = hubbard_square(4, 4, 1.0) if which == "hubbard_mz" else kondo_square(2, 2, 1.0)
    seeds = SEEDS[:3]; nwrap, n_bins, n_sweeps = 5, 2, 2
    model_bin, out_bin = str(tmp_path / "model.bin"), str(tmp_path / "out.bin")
    write_model(model_bin, m, seeds)
    r = subprocess.run([exe, model_bin, str(len(seeds)), str(nwrap), str(n_bins), str(n_sweeps), "1", out_bin], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "acceptance" in r.stdout
    raw = np.fromfile(out_bin, dtype=np.float64)
    obs = raw[: 16 * n_bins].reshape(n_bins, 16); ctl = raw[16 * n_bins: 16 * n_bins + 16]
    rest = raw[16 * n_bins + 16:]; ph = rest[: 2 * len(seeds)].view(np.complex128); G = rest[2 * len(seeds):].view(np.complex128).reshape(m.Ndim, m.Ndim).T
    orcs = []
    for s in seeds:
        o = Oracle(m, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init(); orcs.append(o)
    prev = np.zeros(4)
    for nb in range(n_bins):
        for o in orcs:
            for _ in range(n_sweeps):
                o.sweep(1)
        tot = sum(o.obs() for o in orcs)                  # the oracle accumulates over bins: difference = this bin
        cur = tot - prev; prev = tot
        assert obs[nb, 0] == cur[0] and obs[nb, 1] == cur[1] and abs(obs[nb, 2] - cur[2]) <= 1e-9 * abs(cur[2]) + 1e-9
    for c, o in enumerate(orcs):
        assert abs(ph[c] - o.phase()) < 1e-9
    assert relF(G, orcs[0].green(1)) < TOL_G
    co = [o.control() for o in orcs]
    assert ctl[7] == sum(c["NC_up"] for c in co) and ctl[8] == sum(c["ACC_up"] for c in co) and ctl[11] == 0 and ctl[12] == 0
