// sem2dsolve_b200 -- the reference's `program main` (SRC/main.f90:1-110) for the B200 path:
// reads Par.inp from the working directory, builds the problem in HBM, runs the time loop on the
// device and writes the reference's seismogram and fault files.
//
//   sem2dsolve_b200 [Par.inp] [--precision 4|8] [--device N] [--quiet] [--hash-seed S] [--bench K] [--natural-order]
//
// --energies: what a reference build with COMPUTE_ENERGIES = .true. (SRC/constants.f90:22-27) does: every step, one
// line `time, E_ep, E_k, E_el, E_W` (5D24.16) in energy_sem2d.tab (main.f90:90-93, energy.f90:109-116); E_ep = E_el
// = 0 for the elastic materials of this path (mat_gen.f90:442-445).  Costs one device synchronisation per step.
// --natural-order: row-by-row element order (OPT_RENUMBER = .false.) instead of the reference's default reverse
// Cuthill-McKee order; only the element / node numbering of the grid files and snapshots changes.
//
// --bench K: after init, 5 warm-up steps, then K steps timed on the device (CUDA events, no host traffic); prints
// one JSON line (steps, ms per step, the force kernel's ms per launch, ms per step of every phase) and exits
// without writing files.  Used by bench.py --config for the reference's example decks at scale.
//
// --hash-seed S (S != 0) replaces the homogeneous &MAT_ELASTIC values by the heterogeneous hash medium of the
// synthetic benchmark family (BASELINE.json configs[4]), so that this program can drive it.
//
// Exit code 0 on success; on IO_abort the message is printed as stdio.f90:205-214 does
// ("FATAL ERROR" banner) and the exit code is 1.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sem2d_host.hpp"

using namespace sem2d;

int main(int argc, char** argv) {
  std::string file = "Par.inp";
  bool quiet = false;
  int bench_steps = 0;
  problem_type pb;
  for (int a = 1; a < argc; ++a) {
    const std::string s = argv[a];
    if (s == "--precision" && a + 1 < argc) pb.precision = std::atoi(argv[++a]);
    else if (s == "--device" && a + 1 < argc) pb.device = std::atoi(argv[++a]);
    else if (s == "--quiet") quiet = true;
    else if (s == "--natural-order") pb.renumber = false;
    else if (s == "--energies") pb.compute_energies = true;
    else if (s == "--bench" && a + 1 < argc) bench_steps = std::atoi(argv[++a]);
    else if (s == "--hash-seed" && a + 1 < argc) pb.hash_seed = std::strtoull(argv[++a], nullptr, 10);
    else file = s;
  }
  try {
    const auto t0 = std::chrono::steady_clock::now();
    read_main(pb, file);   // main.f90:27
    init_main(pb);         // main.f90:31
    const auto t1 = std::chrono::steady_clock::now();
    if (!quiet) {
      std::printf("\n Program  S P E C F E M : B200 time-stepping path\n %s\n", pb.title.c_str());
      std::printf("   elements %lld  GLL nodes %lld  ngll %d  ndof %d\n", (long long)pb.nelem_total, (long long)pb.npoin, pb.ngll, pb.ndof);
      std::printf("   Time step (secs)      = %.6E\n   Number of time steps  = %d\n   Total duration (secs) = %.6E\n",
                  pb.time.dt, pb.time.nt, pb.time.total);
      std::printf("   scheme %s\n", pb.time.kind.c_str());
    }
    if (pb.iexec == 0) {  // check mode stops after the set-up (main.f90:66)
      std::printf(" iexec=0: problem checked, not solved\n");
      return 0;
    }
    if (bench_steps > 0) {
      solve(pb, 5);
      float ms = 0.f, kms = 0.f, ph[S2D_NPHASES];
      int64_t l0 = 0, l1 = 0;
      int32_t route = 0;
      s2d_check(pb, s2d_launch_count(pb.gpu, &l0), "bench");
      s2d_check(pb, s2d_time_steps(pb.gpu, bench_steps, &ms), "bench");
      s2d_check(pb, s2d_launch_count(pb.gpu, &l1), "bench");
      s2d_check(pb, s2d_kernel_ms(pb.gpu, &kms), "bench");
      s2d_check(pb, s2d_time_phases(pb.gpu, std::min(bench_steps, 10), ph), "bench");
      s2d_check(pb, s2d_kernel_route(pb.gpu, &route), "bench");
      double vmax = 0, dmax = 0;
      s2d_check(pb, s2d_progress(pb.gpu, &vmax, &dmax), "bench");
      std::printf("{\"npoin\": %lld, \"nelem\": %lld, \"ngll\": %d, \"ndof\": %d, \"scheme\": \"%s\", \"dt\": %.9e, \"steps\": %d, "
                  "\"ms_per_step\": %.6f, \"kernel_ms\": %.6f, \"launches_per_step\": %.2f, \"kv\": %s, \"plastic\": %s, \"vmax\": %.6e, "
                  "\"phases_ms\": [%.6f, %.6f, %.6f, %.6f, %.6f, %.6f, %.6f]}\n",
                  (long long)pb.npoin, (long long)pb.nelem_total, pb.ngll, pb.ndof, pb.time.kind.c_str(), pb.time.dt, bench_steps,
                  ms / bench_steps, kms, (double)(l1 - l0) / bench_steps, pb.has_kv ? "true" : "false", (pb.has_plastic || pb.has_visco || pb.has_damage) ? "true" : "false", vmax, ph[0], ph[1], ph[2],
                  ph[3], ph[4], ph[5], ph[6]);
      return 0;
    }
    // the snapshot files need the grid (spec_grid.f90 writes it at init); PLOT_FIELD at it = 0 (main.f90:38)
    if (snapshot_due(pb, pb.snap_it1) && pb.snap_it1 <= pb.time.nt) SE_write_grid(pb);
    PLOT_FIELD(pb, 0);
    // main.f90:51-99: the loop body runs on the device in chunks that end on the ItInfo lines and on
    // the snapshot steps
    FILE* fen = pb.compute_energies ? std::fopen("energy_sem2d.tab", "w") : nullptr;
    while (pb.it < pb.time.nt) {
      int n = std::min(pb.ItInfo - pb.it % pb.ItInfo, pb.time.nt - pb.it);
      if (fen) n = 1;
      for (int k = 1; k < n; ++k)
        if (snapshot_due(pb, pb.it + k)) {
          n = k;
          break;
        }
      solve(pb, n);
      if (fen) {  // energy_compute + energy_write (main.f90:90-93)
        double Ek = 0, Ew = 0;
        s2d_check(pb, s2d_energy(pb.gpu, &Ek), "energy_compute");
        s2d_check(pb, s2d_energy_w25d(pb.gpu, &Ew), "energy_compute");
        std::fprintf(fen, "%24.16E%24.16E%24.16E%24.16E%24.16E\n", pb.time.time, 0.0, Ek, 0.0, Ew);
      }
      PLOT_FIELD(pb, pb.it);  // main.f90:82
      if (pb.it % pb.ItInfo == 0 && !quiet) {
        double vmax = 0, dmax = 0;
        s2d_check(pb, s2d_progress(pb.gpu, &vmax, &dmax), "main");
        std::printf("Timestep #%8d  t = %11.4E  vmax = %11.4E  dmax = %11.4E\n", pb.it, pb.time.time, vmax, dmax);
      }
    }
    if (fen) std::fclose(fen);
    const auto t2 = std::chrono::steady_clock::now();
    if (pb.rec) {  // main.f90:104
      REC_fetch(pb);
      REC_write(*pb.rec, pb.ndof);
    }
    for (const bc_type& bc : pb.bc)
      if (bc.kind == "DYNFLT") BC_DYNFLT_flush(pb, bc);
    if (!quiet) {
      const double ti = std::chrono::duration<double>(t1 - t0).count(), ts = std::chrono::duration<double>(t2 - t1).count();
      std::printf("\n---  TIME (in seconds) :\n  initialization . . %.3E\n  per timestep . . . %.3E\n  total solver . . . %.3E\n",
                  ti, ts / std::max(pb.time.nt, 1), ts);
      std::printf("  GLL DOF-updates/s  %.3E\n", (double)pb.npoin * pb.ndof * pb.time.nt / ts);
    }
  } catch (const io_abort& e) {  // IO_abort (stdio.f90:205-214)
    std::printf("\n%s\n FATAL ERROR\n %s\n%s\n", std::string(50, '*').c_str(), e.what(), std::string(50, '*').c_str());
    return 1;
  } catch (const std::exception& e) {
    std::printf("\n FATAL ERROR\n %s\n", e.what());
    return 1;
  }
  return 0;
}
