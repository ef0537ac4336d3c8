#!/usr/bin/env python
"""bench.py -- GLL DOF-updates/s of the device-resident SEM2DPACK time loop on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx NX --nz NZ]

Workload (BASELINE.json configs[4]): synthetic Q4 structured mesh, NGLL=5, P-SV (ndof=2),
heterogeneous elastic medium (one coefficient block per element), planar two-sided slip-weakening
fault at mid height, absorbing boundaries on the 4 sides, Ricker point force, 128 receivers, leapfrog,
Courant 0.5, FP64.  One step = one pass of the loop body of SRC/main.f90:51-99.
At N > 1 each rank owns one x-strip of NX element columns (weak scaling, global mesh N*NX x NZ); the same
line then carries a `strong` sub-record (the ONE NX x NZ mesh of BASELINE.json configs[4] split over the N
GPUs) and an `xdev` record (a small global mesh stepped as N strips on N GPUs and as one box on rank 0:
interface copies, fields and fault state compared bit for bit before anything is timed).
The timed state is not at rest: fields start from a seeded random state (s2d_cart_fill_fields).
Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle (a port of the reference's
serial Fortran path: no Fortran compiler exists in this image) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20261017
H = 100.0
NGLL, NDOF, NELAST = 5, 2, 6
W8 = 8
# algorithmic bytes per DOF: K1 = read d, write f, read a, read ibool; step adds the node update
B_K1 = (2 * NDOF * (NGLL - 1) ** 2 * W8 + NELAST * NGLL ** 2 * W8 + 4 * NGLL ** 2) / (NDOF * (NGLL - 1) ** 2)
B_STEP = B_K1 + 6 * W8
B_COEF = NELAST * NGLL ** 2 * W8 / (NDOF * (NGLL - 1) ** 2)   # 37.5 B/DOF of coefficient planes


def moved_bytes_per_dof(fused, store_accel, compact, w=W8, newmark=False):
    """bytes the strip kernel must move per DOF.  Coefficients: all six planes, or (lambda, mu) only in
    the compact mode.  Plain force evaluation: d read, f written.  Fused leapfrog update: d, v read,
    the inverse mass read once per node (w/ndof per DOF), v, d_next (, a) written.  ibool is never read."""
    coef = (2 if compact else NELAST) * NGLL ** 2 * w / (NDOF * (NGLL - 1) ** 2)
    if not fused:
        return coef + 2 * w
    if newmark:   # explicit Newmark also reads a[n-1] and always writes a[n]
        return coef + 4 * w + w / NDOF + 2 * w
    return coef + 4 * w + w / NDOF + (w if store_accel else 0)
METRIC = "GLL DOF-updates/sec"
UNIT = "DOF-updates/s"
CPU_SAMPLE_N = int(os.environ.get("BENCH_CPU_SAMPLE_N", "768"))   # oracle sample mesh (elements per side)
CPU_SAMPLE_STEPS = 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                j = json.load(f)
            if "hbm_gbs" in j:
                return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_deck(nx, nz, nsteps):
    """The same problem as build_engine() as a Par.inp for the oracle (tests/harness.cart_deck)."""
    import harness
    return harness.cart_deck(nx, nz, ngll=NGLL, ndof=NDOF, ezflt=nz // 2, scheme="leapfrog", courant=0.5,
                             nsteps=nsteps, h=H, nrec=0)


def cpu_oracle_rate(nx, nz, nsteps, variant="o3"):
    """DOF-updates/s of the CPU oracle (1 thread; the reference solver is serial) on a bounded sample.
    variant "o3": -O3 -march=x86-64-v3 (BASELINE.md section 4's optimisation level; the .so must run on the GPU
    box's host, so no -march=native); "parity": the -O2 -ffp-contract=off build the parity tests use."""
    import orc
    o = orc.Oracle(synthetic_deck(nx, nz, nsteps + 2), synthetic_seed=SEED, variant=variant)
    ndofs = o.i("npoin") * o.i("ndof")
    o.time_solve(1)
    t = o.time_solve(nsteps)
    o.close()
    return ndofs * nsteps / t, t


def build_engine(nx, nz, rank, world, device, nt_max, precision=8, sync_dt=None, scheme_kind=0, oixd=None):
    from sem2dpack_b200 import CartEngine
    ez = nz // 2
    e = CartEngine(NGLL, NDOF, nx, nz, (rank * nx * H, (rank + 1) * nx * H), (0.0, nz * H), ezflt=ez, seed=SEED,
                   scheme_kind=scheme_kind, courant=0.5, precision=precision, device=device, ix0=rank * nx * (NGLL - 1), iz0=0,
                   halo_left=rank > 0, halo_right=rank < world - 1)
    if sync_dt is not None:
        e.set_dt(sync_dt(e.dt))  # one Courant step for the whole mesh: the minimum over the strips
    xg = world * nx * H
    e.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, xg / 2, 1500.0,
                    oixd=oixd or max(1, (nx * 4 + 1) // 512), oitd=10, nt_max=nt_max)
    sides = [1, 3] + ([4] if rank == 0 else []) + ([2] if rank == world - 1 else [])
    for s in sorted(sides):
        e.add_abso_side(s, False)
    nsrc = 0
    xs, zs = 0.37 * xg, 0.61 * nz * H
    if rank * nx * H <= xs < (rank + 1) * nx * H:
        e.add_force_at(xs, zs, [-0.5, 0.8660254037844386])
        nsrc = 1
    x0, x1 = rank * nx * H, (rank + 1) * nx * H
    e.add_receiver_line(128, (x0 + 0.05 * nx * H, 0.75 * nz * H), (x1 - 0.05 * nx * H, 0.75 * nz * H), "V", 1, nt_max + 1)
    return e, nsrc


FILL = (SEED + 1, 1.0e-3, 1.0)   # seeded non-trivial state of every timed run: |d| <= 1 mm, |v| <= 1 m/s


def attach_halo(e, rank, world, args, torch, dist):
    """x-strip interface exchange of an engine with neighbours; returns the description for `config`"""
    from sem2dpack_b200.strips import attach_halo_exchange, attach_peer_exchange
    halo = "nccl send/recv through torch.distributed"
    ok = 0
    if args.halo == "peer":
        try:
            attach_peer_exchange(e, rank, world)
            ok = 1
        except Exception as ex:  # e.g. CUDA IPC not permitted in this container
            print(f"[bench] rank {rank}: peer-memory halo exchange unavailable ({ex})", file=sys.stderr)
    t = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 1:
        return "peer memory over NVLink (CUDA IPC slots + device flags, no host call on the step path)"
    if args.halo == "peer":
        halo += " (peer-memory mapping failed on some rank)"
    attach_halo_exchange(e, rank, world, precision=args.precision)
    return halo


def xdev_check(rank, world, local, args, torch, dist, sync_dt):
    """Cross-device evidence: a small global mesh stepped as `world` x-strips on `world` GPUs and
    as ONE box on rank 0, same dt / fault / absorbing sides / source / seeded state.  Rank 0 compares the two
    copies of every interface column bit for bit, and every strip and its fault state against the box."""
    import numpy as np
    from sem2dpack_b200.stf import Ricker
    nxs, nzs, k = 48, 40, 30
    nt = k + 8
    e, nsrc = build_engine(nxs, nzs, rank, world, local, nt, args.precision, sync_dt, 0, oixd=1)
    e.commit()
    attach_halo(e, rank, world, args, torch, dist)
    e.fill_fields(*FILL)
    ric = Ricker(2.0, 0.6, 1.0e9)
    e.step(k, ric.table(1, k, e.dt) if nsrc else None)
    LXs, LZs = nxs * (NGLL - 1) + 1, nzs * (NGLL - 1) + 2
    d, v = e.get_window(0, 0, LXs, LZs)
    st = e.fault_state(0, LXs)
    mine = {"d": d, "v": v, "D": st["D"].reshape(NDOF, -1), "V": st["V"].reshape(NDOF, -1), "dt": e.dt}
    e.close()
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    res = None
    if rank == 0:
        from sem2dpack_b200 import CartEngine
        nxg = nxs * world
        g = CartEngine(NGLL, NDOF, nxg, nzs, (0.0, nxg * H), (0.0, nzs * H), ezflt=nzs // 2, seed=SEED, scheme_kind=0,
                       courant=0.5, precision=args.precision, device=local)
        g.set_dt(mine["dt"])
        g.add_fault_swf(0.4, 0.677, 0.525, -120e6, 70e6, 81.6e6, nxg * H / 2, 1500.0, oixd=1, oitd=10, nt_max=nt)
        for sd in (1, 2, 3, 4):
            g.add_abso_side(sd, False)
        g.add_force_at(0.37 * nxg * H, 0.61 * nzs * H, [-0.5, 0.8660254037844386])
        g.commit()
        g.fill_fields(*FILL)
        g.step(k, ric.table(1, k, g.dt))
        LXg = nxg * (NGLL - 1) + 1
        gd, gv = g.get_window(0, 0, LXg, LZs)
        gs = g.fault_state(0, LXg)
        gD, gV = gs["D"].reshape(NDOF, -1), gs["V"].reshape(NDOF, -1)
        g.close()
        iface = all(np.array_equal(allp[r][q][:, :, -1], allp[r + 1][q][:, :, 0]) for r in range(world - 1) for q in "dv")
        worst, fworst = 0.0, 0.0
        for r in range(world):
            x0 = r * nxs * (NGLL - 1)
            for q, ref in (("d", gd), ("v", gv)):
                a, b = allp[r][q], ref[:, :, x0:x0 + LXs]
                worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))
            for q, ref in (("D", gD), ("V", gV)):
                a, b = allp[r][q], ref[:, x0:x0 + LXs]
                fworst = max(fworst, float(np.abs(a - b).max() / max(np.abs(ref).max(), 1e-300)))
        # The two copies of an interface node add the same two addends (own + neighbour's partial sum) and must be
        # bit-identical.  Strips against the one box agree to rounding only: a node shared by four elements is
        # summed as (upper-left + lower-left) + (upper-right + lower-right) across a CTA-group / GPU boundary and
        # as (upper-left + upper-right) + (lower-left + lower-right) inside a group, and the groups fall elsewhere.
        res = {"mesh": f"{nxg}x{nzs} elements as {world} x-strips of {nxs}x{nzs} on {world} GPUs vs one box on rank 0",
               "steps": k, "interface_copies_bitwise_equal": bool(iface), "strips_vs_box_max_rel_diff": worst,
               "fault_state_max_rel_diff": fworst, "tolerance": 1e-11,
               "vmax": float(np.abs(gv).max()), "slip_max": float(np.abs(gD).max()),
               "pass": bool(iface and worst <= 1e-11 and fworst <= 1e-11)}
    flag = torch.tensor([1 if (res is None or res["pass"]) else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if res is not None:
        res["pass_all_reduced"] = bool(int(flag.item()))
    return res


def generic_route_record(n, K, W, local, precision, torch):
    """The drop-in route north_star describes, measured next to the builder's: the SAME n x n heterogeneous box
    (all six planes a(5,5,6) per element -- what matwrk_elast_type%a holds) is stepped (1) as the structured
    builder made it and (2) after its arrays went through the generic C-ABI the way a Fortran host hands them
    over -- s2d_create with ibool in a scrambled (anti-diagonal, RCM-like) element order and the node numbering
    that follows it, s2d_set_elastic with one block per element, rmass in that numbering -- where s2d_commit
    recognises the box from the topology and routes it to the same strip kernel."""
    import numpy as np
    from sem2dpack_b200 import CartEngine, Engine
    e1 = CartEngine(NGLL, NDOF, n, n, (0.0, n * H), (0.0, n * H), seed=SEED, scheme_kind=0, courant=0.5,
                    precision=precision, device=local, coef_mode=1)
    ib, a, rm, _ = e1.get_tables(ibool=True, a=True, rmass=True)
    _, _, hprime = e1.get_gll()
    dt = e1.dt
    e1.commit()
    e1.fill_fields(*FILL)
    e1.step(W, None)
    torch.cuda.synchronize()
    ms1 = e1.time_steps(K)
    k1 = e1.kernel_ms()
    npoin = e1.npoin
    e1.close()
    n2 = NGLL * NGLL
    ib = ib.reshape(n * n, n2)
    ix, iz = np.meshgrid(np.arange(n), np.arange(n))                      # element (ix,iz) at ix + n*iz
    order = np.lexsort((ix.ravel(), (ix + iz).ravel())).astype(np.int64)  # anti-diagonal sweeps, like an RCM front
    ib = ib[order]
    _, first = np.unique(ib.ravel(), return_index=True)                   # nodes renumbered by first occurrence
    old_of_new = ib.ravel()[np.sort(first)]
    new_of_old = np.empty(npoin + 1, np.int32)
    new_of_old[old_of_new] = np.arange(1, npoin + 1, dtype=np.int32)
    ib2 = new_of_old[ib]
    rm2 = rm.reshape(NDOF, npoin)[:, old_of_new - 1]
    a2 = a.reshape(n * n, NELAST * n2)[order]
    del a, ib, rm
    e2 = Engine(NGLL, NDOF, ib2, hprime, rm2, 0, dt, precision=precision, device=local)
    e2.set_elastic(NELAST, a2, np.arange(1, n * n + 1, dtype=np.int32), True)
    del a2
    e2.commit()
    route = e2.route()
    rng = np.random.default_rng(SEED)
    e2.set_fields(FILL[1] * rng.uniform(-1, 1, npoin * NDOF), FILL[2] * rng.uniform(-1, 1, npoin * NDOF))
    e2.step(W, None)
    torch.cuda.synchronize()
    ms2 = e2.time_steps(K)
    k2 = e2.kernel_ms()
    try:   # SURVEY 8d's K1 alone, in the reference's own storage (six planes per element), through the drop-in route
        ms_k1 = e2.time_fint(5)
    except Exception:
        ms_k1 = None
    e2.close()
    ndof = npoin * NDOF
    peak, _ = peaks()
    b = moved_bytes_per_dof(True, False, False, W8 if precision == 8 else 4)
    return {"mesh": f"{n}x{n} elements, heterogeneous, a(5,5,6) per element in HBM, leapfrog, no boundary conditions",
            "builder_value": ndof * K / (ms1 * 1e-3), "generic_value": ndof * K / (ms2 * 1e-3), "unit": UNIT,
            "generic_over_builder": ms1 / ms2, "kernel_route": "strip kernel" if route == 1 else "any-mesh patch kernel",
            "builder_kernel_ms": k1, "generic_kernel_ms": k2, "algorithmic_bytes_per_dof": b,
            "generic_kernel_frac_of_hbm_peak": (b * ndof / (k2 * 1e-3) / 1e9 / peak) if k2 > 0 else None,
            "k1_alone_reference_storage": None if not ms_k1 else {
                "kernel": "k_elem_strip + k_strip_fold, plain force evaluation (compute_Fint) of the routed handle",
                "ms_per_launch": ms_k1, "algorithmic_bytes_per_dof": moved_bytes_per_dof(False, False, False, W8 if precision == 8 else 4),
                "canonical_bytes_per_dof": B_K1, "gdof_per_s": ndof / (ms_k1 * 1e-3) / 1e9,
                "frac": moved_bytes_per_dof(False, False, False, W8 if precision == 8 else 4) * ndof / (ms_k1 * 1e-3) / 1e9 / peak,
                "note": "north_star's K1 target is 60 % of the HBM roofline at the canonical 56.6 B/DOF = 69.4 G DOF/s"},
            "element_order": "anti-diagonal sweeps (ix+iz, ix), node numbering by first occurrence in that order"}


# BASELINE.json configs[0..3]: the reference's own example decks, scaled up so that the device is busy.  Same deck
# text (tests/golden/*.par, copied from EXAMPLES/*/Par.inp), only `nelem` multiplied and the run length replaced.
REF_CONFIGS = {
    "testsh": {"scale": 32, "nelem": (60, 60), "time": ("TotalTime=35.d0", "NbSteps={nt}"),
               "what": "EXAMPLES/TestSH: SH, NGLL=6, homogeneous, ABSORB x2, Ricker force, leapfrog"},
    "lamb": {"scale": 24, "nelem": (40, 20), "time": ("TotalTime=1.5d0, Dt=0.5d-3", "NbSteps={nt}, courant=0.3d0"),
             "what": "EXAMPLES/LambsProblem: P-SV, NGLL=9, free surface + ABSORB x3, Ricker force, leapfrog"},
    "tpv3": {"scale": 24, "nelem": (70, 40), "time": ("TotalTime=16.d0", "NbSteps={nt}"),
             "what": "EXAMPLES/TestFlt2D_SCEC_TPV3_inplane: P-SV, NGLL=6, Kelvin-Voigt layer, one-sided SWF fault, "
                     "ABSORB x2 + DIRNEU, explicit Newmark"},
    "ratestate": {"scale": 12, "nelem": (270, 90), "time": ("TotalTime=4d0", "NbSteps={nt}"),
                  "what": "EXAMPLES/RateState: SH, NGLL=5, one-sided rate-and-state fault (slip law), ABSORB + DIRNEU, "
                          "explicit Newmark"},
    "plastic25d": {"scale": 16, "nelem": (160, 160), "time": ("TotalTime=80", "NbSteps={nt}"),
                   "what": "EXAMPLES/2.5D_plastic: P-SV, NGLL=5, Coulomb plasticity at every GLL point (stateful rheology), "
                           "W = 10 km, SWF + TWF fault, ABSORB x3 + DIRNEU, leapfrog"},
}


def config_bytes_per_dof(ngll, ndof, scheme, kv, w=W8, plastic=False, w25d=False):
    """bytes the engine must move per DOF and step for a builder-made box of that kind: coefficient planes as
    stored (two per GLL point: SH flat planes, or (lambda, mu) of an isotropic P-SV box), fields and inverse mass.
    Fused step (leapfrog / explicit Newmark without KV): d, v in, rmass once per node, v, d_next out (+ a in / out
    for Newmark).  Kelvin-Voigt: the same fused step plus eta per element GLL point (ngll <= 6; v and a are
    double-buffered, the neighbours' velocities come from cache); for ngll > 6 or S2D_KV_FUSED=0 the separate
    predictor pass (d, v, a in; d, v out), force kernel (planes, eta, d, v in; f out), corrector pass (f, rmass, v in;
    a, v out)."""
    n1 = (ngll - 1) ** 2
    coef = 2 * ngll ** 2 * w / (ndof * n1)
    eta = ngll ** 2 * w / (ndof * n1) if kv else 0
    if kv and (ngll > 6 or os.environ.get("S2D_KV_FUSED", "1") == "0"):
        return coef + eta + 13 * w
    # plasticity: the plastic strain (3 per element GLL point) in and out; 2.5D: beta per element GLL point in
    extra = (6 * ngll ** 2 * w / (ndof * n1) if plastic else 0) + (ngll ** 2 * w / (ndof * n1) if w25d else 0)
    return coef + eta + extra + 4 * w + w / ndof + (2 * w if scheme == "newmark" else 0)


def run_ref_config(name, steps, device, scale=None):
    """one of BASELINE.json configs[0..3], scaled, through the host program (sem2dsolve_b200 --bench)"""
    import tempfile
    import harness
    cfg = REF_CONFIGS[name]
    S = scale or cfg["scale"]
    deck = harness.deck(name)
    nx, nz = cfg["nelem"]
    old = f"nelem={nx},{nz}"
    assert old in deck and cfg["time"][0] in deck, name
    deck = deck.replace(old, f"nelem={nx * S},{nz * S}").replace(cfg["time"][0], cfg["time"][1].format(nt=steps + 40))
    exe = os.path.join(ROOT, "sem2dpack_b200", "lib", "sem2dsolve_b200")
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "Par.inp"), "w") as f:
            f.write(deck)
        p = subprocess.run([exe, "--quiet", "--device", str(device), "--bench", str(steps)], cwd=tmp, capture_output=True,
                           text=True, timeout=900)
    if p.returncode != 0:
        return {"config": name, "error": (p.stdout + p.stderr)[-400:]}
    r = json.loads(p.stdout.strip().splitlines()[-1])
    ndofs = r["npoin"] * r["ndof"]
    peak, _ = peaks()
    b = config_bytes_per_dof(r["ngll"], r["ndof"], r["scheme"], r["kv"], plastic=r.get("plastic", False),
                             w25d=("W=" in deck.split("/")[0].replace(" ", "")))
    val = ndofs / (r["ms_per_step"] * 1e-3)
    names = ("predictor", "element_force", "halo_fold_exchange", "sources", "boundary_conditions", "node_update", "outputs")
    return {"config": name, "what": cfg["what"], "mesh": f"{nx * S}x{nz * S} elements (x{S} per side)", "npoin": r["npoin"],
            "ngll": r["ngll"], "ndof": r["ndof"], "scheme": r["scheme"], "kelvin_voigt": r["kv"], "plastic": r.get("plastic", False), "steps": r["steps"],
            "value": val, "unit": UNIT, "ms_per_step": r["ms_per_step"], "launches_per_step": r["launches_per_step"],
            "force_kernel_ms": r["kernel_ms"], "ms_per_step_by_phase": dict(zip(names, r["phases_ms"])),
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_dof": b, "achieved": b * val / 1e9, "peak": peak,
                         "frac": b * val / 1e9 / peak, "unit": "GB/s",
                         "note": "whole step against the bytes the step must move (config_bytes_per_dof)"},
            "vmax": r["vmax"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = CPU_SAMPLE_N
    rates = []
    t0 = time.time()
    for _ in range(max(1, min(args.steps, 3))):
        r, _ = cpu_oracle_rate(n, n, CPU_SAMPLE_STEPS, "o3")
        rates.append(r)
        if time.time() - t0 > 150:
            break
    val = statistics.median(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {args.nx}x{args.nz} Q4, NGLL=5, P-SV heterogeneous + planar fault, leapfrog"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"oracle (C++ port of the serial Fortran path; no Fortran compiler in the image), "
                                   f"{n}x{n}-element sample of the same workload, {CPU_SAMPLE_STEPS} solve() steps x "
                                   f"{len(rates)} repeats, 1 thread (the reference is serial), -O3 -march=x86-64-v3"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--nz", type=int, default=8192)
    ap.add_argument("--precision", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-xdev", action="store_true", help="skip the cross-device parity check (N > 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record (N > 1)")
    ap.add_argument("--config", choices=sorted(REF_CONFIGS), default=None,
                    help="time ONE of the reference's example decks (BASELINE.json configs[0..3]) scaled up, through the "
                         "host program, and print its record instead of the benchmark line")
    ap.add_argument("--config-scale", type=int, default=None, help="elements multiplier per side for --config")
    ap.add_argument("--no-configs", action="store_true", help="skip the reference_configs sub-records (N = 1)")
    ap.add_argument("--generic-n", type=int, default=1536,
                    help="elements per side of the generic-route comparison (N = 1 only; 0 = skip)")
    ap.add_argument("--fint-reps", type=int, default=10)
    ap.add_argument("--coef", choices=["compact", "full"], default="compact",
                    help="coefficient storage: (lambda, mu) per GLL point, or all six planes a(5,5,6) per element")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: every GPU owns an --nx x --nz strip; strong: the --nx x --nz mesh is split over the GPUs")
    ap.add_argument("--scheme", choices=["leapfrog", "newmark"], default="leapfrog",
                    help="time scheme of the workload (BASELINE configs[4]: leapfrog/Newmark); newmark = explicit, beta=0")
    ap.add_argument("--halo", choices=["peer", "nccl"], default="peer",
                    help="x-strip interface exchange: engine kernels writing into the neighbour's memory, or NCCL")
    ap.add_argument("--accel", choices=["last", "every"], default="last",
                    help="accelerations written on the last step of each s2d_step call, or on every step")
    args = ap.parse_args()
    os.environ["S2D_COEF_FULL"] = "1" if args.coef == "full" else "0"
    os.environ["S2D_STORE_ACCEL"] = "1" if args.accel == "every" else "2"
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config:
        print(json.dumps(run_ref_config(args.config, max(args.steps, 5), int(os.environ.get("LOCAL_RANK", "0")), args.config_scale)))
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from sem2dpack_b200 import S2DError
    from sem2dpack_b200.stf import Ricker

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    nt_max = 4 * (K + W) + 64

    def sync_dt(dt_local):
        t = torch.tensor([dt_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # --- cross-device parity on a small global mesh, before anything is timed
    xdev = None
    if world > 1 and not args.no_xdev:
        xdev = xdev_check(rank, world, local, args, torch, dist, sync_dt)
        barrier()

    ric = Ricker(2.0, 0.6, 1.0e9)
    scheme_kind = 1 if args.scheme == "newmark" else 0

    def build(nx, nz_req):
        """the workload on this rank's x-strip, falling back to a shorter mesh if 180 GB cannot hold it"""
        tried = []
        for nzt in [nz_req, (nz_req * 3) // 4, nz_req // 2, nz_req // 4]:
            try:
                e, nsrc = build_engine(nx, nzt, rank, world, local, nt_max, args.precision,
                                       sync_dt if world > 1 else None, scheme_kind)
                e.commit()
                halo = attach_halo(e, rank, world, args, torch, dist) if world > 1 else "none (one strip)"
                e.fill_fields(*FILL)   # the timed region starts from a non-trivial state on every rank
                return e, nsrc, nzt, halo, tried
            except S2DError as ex:
                tried.append(f"{nx}x{nzt}: {ex}")
                torch.cuda.empty_cache()
        raise SystemExit("could not build the workload: " + "; ".join(tried))

    def timed_steps(e, nsrc):
        """W warm-up steps (they also load the stf table the timed replay cycles through), then K steps between
        CUDA events on the engine stream; max over ranks"""
        barrier()
        e.step(W, ric.table(1, W, e.dt) if nsrc else None)
        barrier()
        l0 = e.launch_count()
        clk = ClockSampler(local)
        barrier()
        ms = e.time_steps(K)
        barrier()
        clocks = clk.stop()
        return allmax(ms), e.launch_count() - l0, clocks

    nx, nz = args.nx, args.nz
    if args.scaling == "strong":
        if nx % world:
            raise SystemExit("--scaling strong needs --nx divisible by the number of GPUs")
        nx //= world
    e, nsrc, nz, halo, tried = build(nx, nz)
    ndofs_rank = e.npoin * NDOF
    ms_max, launches, clocks = timed_steps(e, nsrc)
    value = ndofs_rank * world * K / (ms_max * 1e-3)
    # --- dominant kernel: the strip kernel as launched inside the timed steps (CUDA events around every
    # launch on the engine stream), and the plain force stage (strip kernel + halo fold) timed alone
    ms_kernel = e.kernel_ms()
    fused = (os.environ.get("S2D_FUSED", "1") != "0")
    store_accel = args.accel == "every"
    compact = args.coef == "compact"
    w = W8 if args.precision == 8 else 4
    b_moved = moved_bytes_per_dof(fused, store_accel, compact, w, args.scheme == "newmark")
    barrier()
    try:
        ms_fint = e.time_fint(args.fint_reps)
    except Exception as ex:  # the plain evaluation needs a force buffer the fused step does not (--coef full at 8192^2: no room)
        print(f"[bench] plain force evaluation not timed: {ex}", file=sys.stderr)
        ms_fint = float("nan")
    barrier()
    # the O(boundary) kernels are launch-latency bound: reported as time per step, not against the roofline
    phases = e.time_phases(max(3, min(K, 10)))
    barrier()
    peak, peak_src = peaks()
    ach = b_moved * ndofs_rank / (ms_kernel * 1e-3) / 1e9
    b_k1 = moved_bytes_per_dof(False, False, compact, w)   # never claim bytes that are not moved
    ach_k1 = b_k1 * ndofs_rank / (ms_fint * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        key = (f"{nx}x{nz}:f{args.precision * 8}:{'fused' if fused else 'plain'}:{'a' if store_accel else 'noa'}:"
               f"{args.coef}" + (":newmark" if args.scheme == "newmark" else ""))
        traffic = tj.get(key)
    except Exception:
        pass
    # --- end to end through the public API with host buffers: one s2d_step per step with that
    # step's stf row (H2D, through the engine's pinned staging buffer) and a read-back of the step's
    # seismogram row (D2H, likewise)
    it0 = e.it
    n_e2e = max(5, min(K, 20))
    barrier()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        e.step(1, ric.table(it0 + 1 + k, 1, e.dt) if nsrc else None)
        row = e.seis_row(it0 + 1 + k)
    barrier()
    t_e2e = allmax(time.perf_counter() - t0)
    e2e_val = ndofs_rank * world * n_e2e / t_e2e
    h2d = int(allmax(8 * nsrc))          # the rank that owns the source uploads its stf row; max over ranks
    d2h = int(allmax(int(row.nbytes)))
    vmax, dmax = e.progress()
    vmax, dmax = allmax(vmax), allmax(dmax)
    npoin_rank, nelem_rank, dt_run = e.npoin, e.nelem, e.dt
    e.close()
    del e
    torch.cuda.empty_cache()

    # --- strong scaling of the ONE mesh BASELINE.json configs[4] names, split over the GPUs (same line)
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong and args.nx % world == 0:
        es, nsrc_s, nz_s, _, tried_s = build(args.nx // world, args.nz)
        ms_s, launches_s, clocks_s = timed_steps(es, nsrc_s)
        LZg = nz_s * (NGLL - 1) + 2
        ndof_global = (args.nx * (NGLL - 1) + 1) * LZg * NDOF    # unique nodes of the global mesh (interface columns once)
        phases_s = es.time_phases(max(3, min(K, 10)))
        strong = {"workload": f"the ONE {args.nx}x{nz_s} mesh split into {world} x-strips of {args.nx // world}x{nz_s}",
                  "value": ndof_global * K / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / K, "steps": K,
                  "dofs_global": ndof_global, "gpu_launches": int(launches_s), "clocks": clocks_s,
                  "ms_per_step_by_phase": phases_s, "scaling": "strong"}
        es.close()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64" if args.precision == 8 else "f32", "data": "synthetic",
        "config": {"workload": f"synthetic {nx * world}x{nz} Q4 structured mesh ({nx}x{nz} x-strip per GPU), NGLL=5, "
                               "ndof=2 heterogeneous isotropic elastic (material differs at every GLL point) + planar "
                               f"two-sided SWF fault + ABSORB on 4 sides, {args.scheme}, Courant 0.5, 128 receivers/GPU",
                   "coefficients": ("(lambda, mu) per GLL point in HBM, six planes formed in registers" if compact
                                    else "one a(5,5,6) block per element in HBM"),
                   "accel": ("materialised every step" if store_accel else
                             ("written every step (Newmark needs a[n-1])" if args.scheme == "newmark" else
                              "not written by the leapfrog step: formed on demand at s2d_get_fields from one force "
                              "evaluation of d[n] (deferred boundary nodes always keep theirs); round 1 wrote them on "
                              "the last step of every s2d_step call, i.e. every step of the e2e leg")),
                   "initial_state": "seeded random fields on every rank (s2d_cart_fill_fields: |d| <= 1 mm, |v| <= 1 m/s), "
                                    "not the rest state",
                   "halo_exchange": halo, "npoin_per_gpu": npoin_rank, "nelem_per_gpu": nelem_rank, "dt": dt_run,
                   "l2_policy": "working set (>=100 GB per GPU at the default size) far exceeds the 126 MB L2",
                   "requested": f"{args.nx}x{args.nz}", "fallbacks_tried": tried},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": f"k_elem_strip<fused {args.scheme} update>" if fused else "k_elem_strip",
                     "algorithmic_bytes_per_dof": b_moved, "dofs_per_launch": ndofs_rank, "ms_per_launch": ms_kernel,
                     "note": "bytes = what this kernel must move per DOF (coefficients, d, v, rmass in; v, d_next(, a) out); "
                             "SURVEY 8d's canonical K1+update figure is %.1f B/DOF (%.1f with a stored)" % (B_STEP, B_STEP + W8),
                     "k1_alone": None if ms_fint != ms_fint else
                                 {"kernel": "k_elem_strip + k_strip_fold, plain force evaluation", "ms_per_launch": ms_fint,
                                  "algorithmic_bytes_per_dof": b_k1, "achieved": ach_k1, "frac": ach_k1 / peak,
                                  "canonical_bytes_per_dof": B_K1, "gdof_per_s": ndofs_rank / (ms_fint * 1e-3) / 1e9,
                                  "note": "BASELINE.md's K1 target (60 % of the roofline at the canonical 56.6 B/DOF) is "
                                          "69.4 G DOF/s; frac is quoted on the bytes this kernel really moves"},
                     "full_step": {"algorithmic_bytes_per_dof": b_moved, "canonical_bytes_per_dof": B_STEP,
                                   "achieved": b_moved * value / world / 1e9, "frac": b_moved * value / world / 1e9 / peak,
                                   "note": "whole step (strip kernel + fold + boundary + deferred-node kernels) against "
                                           "the bytes the fused kernel must move"}},
        "ms_per_step_by_phase": dict(phases, note="CUDA events between the phases of a step on the engine stream; sources, "
                                                  "boundary conditions (4 ABSORB sides + DYNFLT), deferred nodes and outputs "
                                                  "are O(boundary) and launch-latency bound (SURVEY 8d)"),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e},
        "gpu_launches": int(launches), "clocks": clocks,
        "check": {"vmax": vmax, "dmax": dmax, "note": "max over all ranks of max|v|, max|d| after the run"},
    }
    def guarded(what, fn, *a):   # a sub-record that fails must not take the headline line with it
        try:
            return fn(*a)
        except Exception as ex:
            print(f"[bench] {what} failed: {ex}", file=sys.stderr)
            return {"record": what, "error": str(ex)[-300:]}
    if world == 1 and args.generic_n > 0:
        line["generic_route"] = guarded("generic_route", generic_route_record, min(args.generic_n, args.nx), K, W, local,
                                        args.precision, torch)
    if world == 1 and not args.no_configs:   # BASELINE.json configs[0..3] at scale, same run (records, not the headline)
        line["reference_configs"] = [guarded(c, run_ref_config, c, max(10, min(K, 30)), local)
                                     for c in ("testsh", "lamb", "tpv3", "ratestate", "plastic25d")]
    if strong is not None:
        line["strong"] = strong
    if xdev is not None:
        line["xdev"] = xdev
    def cpu_record():
        r, tcpu = cpu_oracle_rate(CPU_SAMPLE_N, CPU_SAMPLE_N, CPU_SAMPLE_STEPS, "o3")
        r2, tcpu2 = cpu_oracle_rate(CPU_SAMPLE_N, CPU_SAMPLE_N, max(2, CPU_SAMPLE_STEPS // 2), "parity")
        return {"value": r, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"oracle (C++ port of the serial Fortran path), {CPU_SAMPLE_N}x{CPU_SAMPLE_N}"
                                          f"-element sample of the same workload, {CPU_SAMPLE_STEPS} solve() steps, "
                                          f"{tcpu:.1f} s, 1 thread; of {os.cpu_count()} host cores",
                                "build": "-O3 -march=x86-64-v3 (BASELINE.md section 4)",
                                "parity_build_value": r2,
                                "parity_build": "-O2 -ffp-contract=off (the build the parity tests compare against)"}
    if rank == 0 and not args.no_cpu:
        line["cpu_baseline"] = guarded("cpu_baseline", cpu_record)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
