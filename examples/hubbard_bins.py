#!/usr/bin/env python
"""End-to-end example of the batched mode: N_bin bins of N_sweep sweeps of many Hubbard chains on one GPU with the device-side
measurements switched on, the per-bin reduction over chains (and over ranks under torchrun: alf_b200.parallel.reduce_sum replaces
MPI_REDUCE) and ALF's text bin files (`Part_scal`, `Green_eq`, `SpinZ_eq`, `Den_eq`, `Green_tau`, `SpinZ_tau`, `Den_tau`, ...) that the
unchanged Analysis tools read.

    python examples/hubbard_bins.py --L1 4 --L2 4 --beta 2 --chains 32 --bins 4 --sweeps 10 --out /tmp/run
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alf_b200.api import AlfB200  # noqa: E402
from alf_b200.bins import print_bin_latt, print_bin_vec  # noqa: E402
from alf_b200.model import hubbard_square, z2_matter_square  # noqa: E402
from alf_b200 import conf  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--L1", type=int, default=4); ap.add_argument("--L2", type=int, default=4)
    ap.add_argument("--beta", type=float, default=2.0); ap.add_argument("--dtau", type=float, default=0.1); ap.add_argument("--U", type=float, default=4.0)
    ap.add_argument("--chains", type=int, default=32); ap.add_argument("--nwrap", type=int, default=10)
    ap.add_argument("--bins", type=int, default=2); ap.add_argument("--sweeps", type=int, default=5); ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--ltau", type=int, default=1); ap.add_argument("--out", default="."); ap.add_argument("--seed0", type=int, default=4711)
    ap.add_argument("--model", default="hubbard", choices=["hubbard", "hubbard_continuous", "z2_matter"],
                    help="hubbard_continuous: Continuous = .true. (type-3 fields); z2_matter: Hamiltonian_Z2_Matter with S0 / Global_move_tau tables on the device")
    ap.add_argument("--restart", action="store_true", help="start from confin_<chain> in --out instead of a random configuration (Fields_in)")
    a = ap.parse_args(argv)
    os.makedirs(a.out, exist_ok=True)
    if a.model == "z2_matter":
        model = z2_matter_square(a.L1, a.L2, beta=a.beta, dtau=a.dtau)
    else:
        model = hubbard_square(a.L1, a.L2, beta=a.beta, dtau=a.dtau, U=a.U, continuous=(a.model == "hubbard_continuous"))
    g = AlfB200(model, n_chains=a.chains, nwrap=a.nwrap)
    g.set_seeds([a.seed0 + 7919 * c for c in range(a.chains)]); g.fields_set()
    if a.restart:
        conf.read_confs(g, a.out)                                  # confin_<chain>: random-number state and nsigma%f of every chain
    g.init_sweep()
    g.sweep(a.warmup, 0)
    g.obs_eq_enable(True)
    if a.ltau:
        g.obs_tau_enable(True)
    names = ["Green", "SpinZ", "SpinXY", "Den"]

    def background(nm, bg):      # Obs_Latt0 is accumulated for SpinZ (ObsZ) and Den only (Prog/Predefined_Obs_mod.F90:190-191,514-515); Green and SpinXY: 0
        return bg[0].sum(0) if nm == "SpinZ" else bg[1].sum(0) if nm == "Den" else np.zeros(model.n_orb)
    unit = dict(n_coord=model.n_coord, orb_pos=model.orb_pos)
    for nb in range(a.bins):
        g.obs_reset()
        if a.ltau:
            g.obs_tau_reset()
        acc0, bg0, n0, s0 = g.obs_eq()                          # equal-time accumulators are cumulative: difference per bin
        g.sweep(a.sweeps, a.ltau)
        ob = g.obs()                                             # [N_meas (chain-slices), sum sign, Re/Im sum Part]
        print_bin_vec(os.path.join(a.out, "Part_scal"), [complex(ob[2], ob[3])], ob[1], ob[0] / a.chains, a.chains, description=["Particle number"])
        acc, bg, n, s = g.obs_eq()
        acc, bg, n, s = acc - acc0, bg - bg0, n - n0, s - s0
        for ch, nm in enumerate(names):
            print_bin_latt(os.path.join(a.out, nm), acc[ch].transpose(0, 2, 1, 3), background(nm, bg),
                           s, n / a.chains, a.chains, model.latt, channel="---", **unit)
        if a.ltau:
            acc, bg, n, s = g.obs_tau()
            for ch, nm in enumerate(names):
                print_bin_latt(os.path.join(a.out, nm), acc[ch].transpose(0, 2, 1, 3), background(nm, bg),
                               s, n / a.chains, a.chains, model.latt, dtau=a.dtau, channel="P", **unit)
        c = g.control()
        print(f"bin {nb}: acceptance {c['ACC_up'] / max(c['NC_up'], 1):.3f}, precision Green max {c['XMAXG']:.2e}, <sign> {ob[1] / ob[0]:.3f}, <N> {ob[2] / ob[0]:.4f}")
    for f in conf.write_confs(g, a.out):                           # confout_<chain>, then renamed as ALF's out_to_in.sh does
        os.replace(f, os.path.join(os.path.dirname(f), os.path.basename(f).replace("confout", "confin")))
    g.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
