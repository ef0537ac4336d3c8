"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only;
/root/reference does not exist on the GPU box).

  * *.par       -- the four CPU-runnable example decks of BASELINE.json and four more MESH_CART examples
                   (EXAMPLES/*/Par.inp with comment lines dropped); they are INPUT data, the namelist text
                   the reference reads.
  * refdata.npz -- the reference's own known-answer artefacts for those decks:
                   TestSH/uyref.mat (analytic SH trace used by EXAMPLES/TestSH/analyze_test.m),
                   LambsProblem/U{x,z}_file_ascii (EX2DDIR traces, analyze_test.m) and the misfits
                   recorded in LambsProblem/test.out:12, RateState/{Tau,Ux,Vx}_{0,3,6,9}km_ascii.
"""
import os
import re

import numpy as np
import scipy.io

REF = "/root/reference/EXAMPLES"
HERE = os.path.dirname(os.path.abspath(__file__))
DECKS = {"testsh": "TestSH", "lamb": "LambsProblem", "tpv3": "TestFlt2D_SCEC_TPV3_inplane", "ratestate": "RateState",
         # round 2: the time-solver lines of InaBox/info:224-225 pin dt / nt; the others are the decks VERDICT r1 names
         "inabox": "InaBox", "velweak": "Velocity_weakening", "inplane25d": "2.5D_inplane", "kvfz": "Kelvin_Visco_FZ",
         # SURVEY 8f(4), first stateful rheology: Coulomb plasticity
         "plastic25d": "2.5D_plastic/psi_45_S_0.56_CF_0.63_W_10", "attenuation": "Attenuation", "damage": "Damage"}


def main():
    for short, d in DECKS.items():
        with open(os.path.join(REF, d, "Par.inp")) as f:
            lines = [ln.rstrip() for ln in f if ln.strip() and not ln.lstrip().startswith("#")]
        with open(os.path.join(HERE, short + ".par"), "w") as f:
            f.write("\n".join(lines) + "\n")
    out = {}
    m = scipy.io.loadmat(os.path.join(REF, "TestSH", "uyref.mat"))
    out["testsh_uref"] = m["uref"].ravel().astype(np.float64)
    out["testsh_tref"] = m["tref"].ravel().astype(np.float64)
    for c in ("x", "z"):
        out[f"lamb_u{c}"] = np.loadtxt(os.path.join(REF, "LambsProblem", f"U{c}_file_ascii")).astype(np.float64)
    with open(os.path.join(REF, "LambsProblem", "test.out")) as f:
        txt = f.read()
    out["lamb_misfits"] = np.array([float(x) for x in re.search(r"uz2 = ([0-9.\s]+)", txt).group(1).split()])
    for q in ("Tau", "Ux", "Vx"):
        for km in (0, 3, 6, 9):
            out[f"ratestate_{q}_{km}km"] = np.loadtxt(os.path.join(REF, "RateState", f"{q}_{km}km_ascii"))
    np.savez_compressed(os.path.join(HERE, "refdata.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
