"""Source time functions evaluated on the host, one scalar per source per step, exactly where the
reference evaluates them (SO_add -> STF_get, SRC/src_gen.f90:300-303)."""
import math

import numpy as np


class Ricker:
    """stf_ricker_type (SRC/stf_ricker.f90:12-15, :89-101).  f0, onset and ampli are read by the
    reference as single-precision `real` and widened (stf_ricker.f90:53,71-73); the same rounding
    is applied here."""

    def __init__(self, f0, onset, ampli=1.0):
        self.f0 = float(np.float32(f0))
        self.t0 = float(np.float32(onset))
        self.ampli = float(np.float32(ampli))

    def __call__(self, t):
        arg = math.pi * self.f0 * (t - self.t0)
        arg = arg * arg
        return -self.ampli * (1.0 - 2.0 * arg) * math.exp(-arg)

    def table(self, it_first, nsteps, dt, tdelay=0.0, scale=1.0):
        """rows it_first .. it_first+nsteps-1 of the amplitude table s2d_step consumes (time = it*dt,
        SRC/main.f90:53)."""
        return np.array([[self((it_first + k) * dt - tdelay) * scale] for k in range(nsteps)], dtype=np.float64)
