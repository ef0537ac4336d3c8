/* sem2d_b200.h -- C-ABI of the B200-native SEM2DPACK time-stepping engine.
 *
 * This is the drop-in boundary for ONE path of SEM2DPACK (jpampuero/sem2dpack): the explicit
 * time step `solve(pb)` (SRC/solver.f90:20-84,140-160) with everything it calls per step --
 * compute_Fint / MAT_Fint / MAT_ELAST_f (SRC/solver.f90:273-320, SRC/mat_gen.f90:418-460,
 * SRC/mat_elastic.f90:396-799), SO_add (SRC/src_gen.f90:290-317), BC_apply
 * (SRC/bc_gen.f90:256-308), REC_store (SRC/receivers.f90:309-344) and BC_write
 * (SRC/bc_gen.f90:313-337) -- kept resident on one GPU.  The reference has no FFI layer; its
 * boundary is the Fortran module API of one executable.  Each entry point below names the
 * Fortran interface (file:line under SRC/) whose data it receives or whose work it replaces;
 * INTEGRATION.md shows the ISO_C_BINDING stubs that bind them.
 *
 * Conventions (the reference's own, so the Fortran host passes its arrays untouched):
 *   - every pointer is a HOST pointer unless the name says `_dev`; data are copied at call time and
 *     the caller keeps ownership;
 *   - arrays are column-major, indices (nodes, elements, boundary nodes) are 1-based, integers
 *     are 32-bit, reals are IEEE double regardless of the engine's compute precision;
 *   - fields are (npoin, ndof): component c of node k at [k-1 + npoin*c];
 *   - every function returns 0 on success or a negative S2D_E* code; s2d_last_error() gives the
 *     text the shim passes to IO_abort (SRC/stdio.f90:205-214);
 *   - one handle per process / GPU; calls on one handle are not re-entrant across threads.
 * There is no CPU fallback: if no CUDA device is usable s2d_create fails with S2D_ENODEV.
 */
#ifndef SEM2D_B200_H
#define SEM2D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2D_OK 0
#define S2D_EINVAL (-1)   /* bad argument */
#define S2D_ENODEV (-2)   /* no usable CUDA device */
#define S2D_ECUDA (-3)    /* CUDA runtime error (text in s2d_last_error) */
#define S2D_ESTATE (-4)   /* call out of order (e.g. step before commit) */
#define S2D_ESOLVER (-5)  /* device-side abort: NR_Solver exceeded 200 iterations (bc_dynflt_rsf.f90:463-466) */
#define S2D_ENOMEM (-6)

typedef struct s2d_engine* s2d_handle;

/* timescheme_type (SRC/time.f90:5-11).  kind:
 *   0 'leapfrog'  (solver.f90:140-160)      1 'newmark' (solver.f90:42-84)
 *   2 'HHT-alpha' (solver.f90:89-128): beta, gamma, alpha as TIME_read derives them (time.f90:232-246)
 *   3 symplectic  (solver.f90:169-199; 'symp_PV', 'symp_PFR', 'symp_PEFRL', 'symp_PEFRL4'...):
 *     nstages and the coefficient tables time%a(1:nstages+1), time%b(1:nstages) (time.f90:248-300);
 *     as in the reference, boundary conditions are not applied by this scheme.
 * The source amplitudes s2d_step receives are evaluated by the host at the times the scheme uses:
 * it*dt for kinds 0/1, it*dt+(alpha-1)*dt for kind 2, and one row per STAGE for kind 3
 * (t = (it-1)*dt + dt*sum(a(1:k)), solver.f90:186-191). */
#define S2D_MAX_STAGES 8
typedef struct {
  int32_t kind;
  double dt, beta, gamma, alpha;
  int32_t nstages;
  double coa[S2D_MAX_STAGES + 1], cob[S2D_MAX_STAGES];
} s2d_scheme;

/* element-force / assembly kernel variants */
#define S2D_ASM_PATCH 0    /* default: CTA-patch kernel, in-patch colours + ordered halo sums (deterministic) */
#define S2D_ASM_COLOR 1    /* one launch per greedy mesh colour (deterministic) */
#define S2D_ASM_ATOMIC 2   /* one launch, atomicAdd assembly (measured alternative, not deterministic) */

/* ---- life cycle ------------------------------------------------------------------------- */

/* Receives what init_main (SRC/init.f90:16-131) has built: sem_grid_type%ibool(ngll,ngll,nelem)
 * and %hprime (SRC/spec_grid.f90:49-64), the inverted mass `rmass` (init.f90:112-116), the time
 * scheme.  precision = 8 (FP64, parity mode) or 4 (FP32 fields and coefficients).
 * device < 0 selects the current CUDA device. */
int s2d_create(s2d_handle* h, int32_t ngll, int32_t ndof, int32_t nelem, int32_t npoin,
               const int32_t* ibool, const double* hprime, const double* rmass, int32_t precision,
               const s2d_scheme* scheme, int32_t device);
int s2d_destroy(s2d_handle h);
const char* s2d_last_error(s2d_handle h);
const char* s2d_version(void);

/* matwrk_elast_type%a (SRC/mat_elastic.f90:11-14,290-360): ncoefsets blocks a(ngll,ngll,nelast),
 * elem2set(nelem) 1-based block of each element (mat_gen.f90:357-365 shares one block between
 * homogeneous elements).  nelast = 2|3 (SH flat|general), 6|10 (P-SV flat|general).
 * beta25d: matwrk_elast_type%beta(ngll,ngll) per coefficient block (MAT_ELAST_init_25D, mat_elastic.f90:363-383),
 * the finite-seismogenic-width term of 2.5D runs (&GENERAL W): every element force gets
 * - beta * d (MAT_ELAST_add_25D_f, :447-459, called from MAT_Fint, mat_gen.f90:440, on the Kelvin-Voigt-modified
 * d when the element carries KV); NULL when W is infinite.
 * kd2 != 0 selects the ELAST_KD2_* form the reference uses when ngll == OPT_NGLL
 * (mat_elastic.f90:412-423,612), 0 the ELAST_KD1_* form (:489). */
int s2d_set_elastic(s2d_handle h, int32_t nelast, int32_t ncoefsets, const double* a,
                    const int32_t* elem2set, const double* beta25d, int32_t kd2);

/* matwrk_kv_type%eta (SRC/mat_kelvin_voigt.f90:117-150): eta(ngll,ngll,nkv), already times dt
 * when ETAxDT; elem_ids(nkv) 1-based elements carrying it. */
int s2d_set_kv(s2d_handle h, int32_t nkv, const int32_t* elem_ids, const double* eta);

/* assembled nodal mass (SRC/mat_mass.f90:29-61, column 1) before BC_init touched it; only needed
 * for s2d_energy (SRC/energy.f90:49-106). */
int s2d_set_mass(s2d_handle h, const double* mass);

/* bc_abso_type (SRC/bc_abso.f90:38-46), built by BC_ABSO_init (:115-266): node(np) bulk nodes,
 * C(np,ndof); is_flat=0 needs n(np,2); stacey!=0 needs bibool(ngll,nbe) and K(ngll,2,nbe). */
int s2d_add_abso(s2d_handle h, int32_t np, const int32_t* node, const double* C, int32_t is_flat,
                 const double* n, int32_t stacey, int32_t nbe, const int32_t* bibool,
                 const double* K);

/* bc_dirneu_type (SRC/bc_dirneu.f90:17-23): kind 1 = Neumann, 2 = Dirichlet per component.
 * A Neumann component with a source time function adds stf(t)*B(np) (:148-169); its amplitude
 * comes per step through s2d_step's `bc_ampli` (NULL B = homogeneous Neumann, a no-op). */
int s2d_add_dirneu(s2d_handle h, int32_t np, const int32_t* node, int32_t kind_h, int32_t kind_v,
                   const double* B_h, const double* B_v);

/* bc_periodic_type (SRC/bc_periodic.f90:11-14): f(master) += f(slave); f(slave) = f(master)
 * (BC_PERIO_set_field, :77-86), applied before every other boundary (bc_gen.f90:271-281).  The mass
 * handed to s2d_create already carries the same sum (BC_PERIO_init, :73). */
int s2d_add_periodic(s2d_handle h, int32_t np, const int32_t* master, const int32_t* slave);

/* bc_dynflt_type (SRC/bc_dynflt.f90:18-38) after BC_DYNFLT_init (:231-520). */
typedef struct {
  int32_t np;
  const int32_t* node1;    /* (np) */
  const int32_t* node2;    /* (np) or NULL: one-sided fault, tags(2)=0 (:700-716) */
  const double* n1;        /* (np,2) */
  const double* B;         /* (np,ndof) */
  const double* invM1;     /* (np,ndof) */
  const double* invM2;     /* (np,ndof) or NULL */
  const double* Z;         /* (np,ndof) */
  const double* T0;        /* (np,2) */
  const double* cohesion;  /* (np) */
  const double* coord;     /* (2,np) */
  const double* V0;        /* (np,ndof) initial bc%V (RSF: &BC_DYNFLT V) or NULL = 0 */
  double CoefA2V, CoefA2D; /* time.f90:426-456 */
  int32_t allow_opening;
  /* slip weakening, swf_type (bc_dynflt_swf.f90:12-20); swf_kind 0 = absent */
  int32_t swf_kind, swf_healing;
  const double *swf_dc, *swf_mus, *swf_mud, *swf_p, *swf_alpha, *swf_theta; /* (np) each */
  /* rate and state, rsf_type (bc_dynflt_rsf.f90:14-22); rsf_kind 0 = absent */
  int32_t rsf_kind;
  const double *rsf_dc, *rsf_mus, *rsf_a, *rsf_b, *rsf_Vstar, *rsf_theta, *rsf_Vc; /* (np) each */
  /* time weakening, twf_type (bc_dynflt_twf.f90:12-16); twf_kind 0 = absent */
  int32_t twf_kind;
  double twf_X, twf_Z, twf_mus, twf_mud, twf_mu0, twf_L, twf_V, twf_T, twf_Dc;
  /* normal stress response, normal_type (bc_dynflt_normal.f90:8-13), kinds 0..3 */
  int32_t normal_kind;
  double normal_T, normal_L, normal_V;
  /* outputs (bc_dynflt.f90:455-458): records of nodes oix1:oixn:oixd every oitd steps from step
   * oit on; nt_max = number of steps the history buffers must hold (time%nt) */
  int32_t oix1, oixn, oixd, oit, oitd, nt_max;
} s2d_dynflt_desc;
int s2d_add_dynflt(s2d_handle h, const s2d_dynflt_desc* desc, int32_t* fault_id);

/* src_force_type (SRC/src_force.f90:77-90): f(iglob,:) += dir(:)*ampli(t); returns source index. */
int s2d_add_force(s2d_handle h, int32_t iglob, const double dir[2], int32_t* src_id);
/* so_moment_type after SRC_MOMENT_init (SRC/src_moment.f90:129-180): the terms of SRC_MOMENT_add
 * (:183-197) in the order that routine applies them -- for every element k that holds the source node,
 * iglob_xi(:,k) with coef_xi(:,:,k), then iglob_eta(:,k) with coef_eta(:,:,k):
 *   f(node[t], c) += ampli(t_step) * coef[t + nterms*c],  t = 0..nterms-1 in sequence.
 * Shares the source index space (columns of src_ampli) with s2d_add_force. */
int s2d_add_moment(s2d_handle h, int32_t nterms, const int32_t* node, const double* coef, int32_t* src_id);

/* rec_type (SRC/receivers.f90:9-20): field 'D','V' or 'A'; nt_rec = time%nt/isamp+1 (:188);
 * at_node: iglob(nx); else einterp(nx) + interp(ngll*ngll,nx) (:231-303). */
int s2d_add_receivers(s2d_handle h, int32_t nx, char field, int32_t isamp, int32_t nt_rec,
                      int32_t at_node, const int32_t* iglob, const int32_t* einterp,
                      const double* interp);

/* Freezes the configuration: builds the colouring / patch plan, uploads tables, writes the it=0
 * fault record (bc_gen.f90:249) and the it=0 seismogram sample (main.f90:35). */
int s2d_commit(s2d_handle h, int32_t assembly_variant);

/* ---- time loop -------------------------------------------------------------------------- */

/* FIELDS (SRC/fields.f90:7-11); any pointer may be NULL.  Only for init / snapshots / exit. */
int s2d_set_fields(s2d_handle h, const double* displ, const double* veloc, const double* accel);
int s2d_get_fields(s2d_handle h, double* displ, double* veloc, double* accel);

/* nsteps iterations of the body of main.f90:51-99: it=it+1, time=it*dt, solve, REC_store,
 * BC_write -- without host traffic.  src_ampli[nsteps][nsrc] = STF_get(t-tdelay)*ampli per source
 * (src_gen.f90:300-303), bc_ampli[nsteps][2*ndirneu] = Neumann stf values; NULL when none. */
int s2d_step(s2d_handle h, int32_t nsteps, const double* src_ampli, const double* bc_ampli);

/* One compute_Fint (solver.f90:273-320) on the current displ/veloc, no time integration:
 * fint(npoin,ndof) = -K d.  For kernel parity tests. */
int s2d_compute_fint(s2d_handle h, double* fint);

int s2d_get_it(s2d_handle h, int32_t* it);

/* REC_write's payload (receivers.f90:351-392): sis(nt_rec,nx,ndof) float32. */
int s2d_get_seis(s2d_handle h, float* sis);

/* one sample of every trace: row(nx,ndof) float32 = sis(it/isamp+1,:,:); it must be a sampled step. */
int s2d_get_seis_row(s2d_handle h, int32_t it, float* row);

/* BC_DYNFLT_write's payload (bc_dynflt.f90:751-778): records[nout][6][onx] float32 in file order
 * (D, V, T1, T2, MU, Tstick); potency[ncalls][2*(ndof+1)] (one line of FltXX_potency_sem2d.tab
 * per BC_write call, the it=0 call included).  Either pointer may be NULL; counts always set. */
int s2d_get_fault(s2d_handle h, int32_t fault_id, float* records, int32_t* nout, double* potency,
                  int32_t* ncalls);
/* current FP64 fault state, (np,*) column-major: D,V (np,ndof); T,Tstick (np,2); MU,theta,sigma (np) */
int s2d_get_fault_state(s2d_handle h, int32_t fault_id, double* D, double* V, double* T,
                        double* Tstick, double* MU, double* theta, double* sigma);

/* main.f90:73-76 progress line: maxval|veloc|, maxval|displ|. */
int s2d_progress(s2d_handle h, double* vmax, double* dmax);
/* energy.f90:49-106 kinetic energy 0.5*sum(M v.v) (needs s2d_set_mass). */
int s2d_energy(s2d_handle h, double* E_k);
/* energy.f90:84-104: E_W = 1/2 sum(beta * d.d), the additional elastic energy of a 2.5D run (0 when W is infinite). */
int s2d_energy_w25d(s2d_handle h, double* E_W);

/* ---- plan introspection (parity of the colouring) ------------------------------ */
/* greedy first-fit colours (ascending element id, conflict = shares a GLL node), 0-based, as used
 * by S2D_ASM_COLOR; computed on the device side of the library.  color(nelem). */
int s2d_get_coloring(s2d_handle h, int32_t* ncolors, int32_t* color);

/* ---- kernel routing ---------------------------------------------------------------------- */
/* s2d_commit looks at the topology of ibool: a MESH_CART box (CART_build, mesh_cartesian.f90:219-314) in ANY
 * element order -- in particular the RCM order the reference uses by default (OPT_RENUMBER, constants.f90:11;
 * MESH_STRUCTURED_renumber, mesh_structured.f90:204-269) -- with or without the split-node row of `ezflt`, flat
 * coefficient planes (nelast 2 | 6) and the OPT_NGLL choice of kd2, is moved onto the GLL lattice and runs on the
 * z-marching strip kernel, exactly like a builder-made box; everything else keeps the any-mesh kernels.  The API
 * keeps the caller's node and element numbering either way.  S2D_ROUTE_STRIP=0 (environment) disables the routing.
 * route: 0 = any-mesh kernel (patch / colour / atomic variant), 1 = strip kernel. */
int s2d_kernel_route(s2d_handle h, int32_t* route);
/* The recognition on its own (host only, no device needed): nx = nz = 0 when ibool is not such a box; otherwise the
 * box, the number of element rows below the split-node row (ezflt, 0 = none), the position (ex, ez)(nelem) of
 * every element and the GLL lattice position (gx, gz)(npoin) of every node, 0-based; gz counts the duplicated
 * fault row.  lower_hint: a node of the lower block (node1 of a two-sided fault) or 0.  Pointers may be NULL. */
int s2d_detect_structured(int32_t ngll, int32_t nelem, int32_t npoin, const int32_t* ibool, int32_t lower_hint,
                          int32_t* nx, int32_t* nz, int32_t* ezflt, int32_t* ex, int32_t* ez, int32_t* gx, int32_t* gz);

/* MESH_STRUCTURED_renumber (SRC/mesh_structured.f90:204-269) + genrcm (SRC/rcm.f90): the reverse Cuthill-McKee
 * order the reference gives the elements of every structured mesh by default (OPT_RENUMBER, constants.f90:10-15).
 * perm(nx*nz), 1-based: perm(new) = old, old = i + nx*(j-1) for element (i,j) of CART_build.  Host only. */
int s2d_rcm_box(int32_t nx, int32_t nz, int32_t* perm);

/* ---- measurement hooks ------------------------------------------------------------------- */
/* Times `reps` launches of the element-force+assembly stage alone (CUDA events on the engine's
 * stream), fields untouched apart from accel; returns average milliseconds per launch. */
int s2d_time_fint(s2d_handle h, int32_t reps, float* ms_avg);
/* Times nsteps full steps on the engine stream with CUDA events (no host traffic inside). */
int s2d_time_steps(s2d_handle h, int32_t nsteps, float* ms_total);
/* Average duration (ms per launch, CUDA events on the engine's stream) of the dominant kernel --
 * the strip kernel of a builder-made engine -- over the launches of the last s2d_time_steps call. */
int s2d_kernel_ms(s2d_handle h, float* ms);
/* Runs nsteps steps with CUDA events between the phases of a step on the engine's stream and returns the average
 * milliseconds per step of each: ms_phase[S2D_NPHASES] = [0] step counter / predictor, [1] element-force kernel
 * (fused with the node update where it is), [2] halo folds and x-strip interface exchange, [3] SO_add,
 * [4] BC_apply (absorbing, Dirichlet / Neumann, dynamic faults), [5] node update (deferred nodes or corrector),
 * [6] REC_store + BC_write.  SURVEY 8d: the O(boundary) kernels are reported as time per step. */
#define S2D_NPHASES 7
int s2d_time_phases(s2d_handle h, int32_t nsteps, float* ms_phase);
/* number of kernels the engine has launched so far */
int s2d_launch_count(s2d_handle h, int64_t* n);
/* raw CUDA stream (cudaStream_t) the engine launches on, for external event timing */
int s2d_stream(s2d_handle h, void** stream);

/* Coulomb plasticity (kind='PLAST': MAT_PLAST_read / MAT_PLAST_init_elem_work / MAT_PLAST_stress, mat_plastic.f90:66-118,
 * 148-218,281-387; MAT_Fint's strain -> stress -> force branch, mat_gen.f90:445-449, with MAT_strain_PSV :752-775 and
 * MAT_forces :834-866).  par(6,nsets) = coh, phi [degrees], Tv, e0(3) of every plastic material as &MAT_PLASTIC gives
 * them; elem_set(nelem), natural element order: 0 = elastic element, k = plastic material k (1-based).  lambda, mu come
 * from s2d_cart_set_material's rho, cp, cs (uniform inside a plastic element, as in the reference).  The plastic strain
 * ep(ngll,ngll,3) of every element lives on the device and is advanced by every force evaluation (update = .true.);
 * s2d_cart_get_plastic_strain returns it as (ngll,ngll,3,nelem), natural element order (MAT_PLAST_export).
 * P-SV, ngll <= 6, at most 7 plastic materials, no Kelvin-Voigt elements in the same problem. */
int s2d_cart_set_plastic(s2d_handle h, int32_t nsets, const double* par, const int32_t* elem_set);
int s2d_cart_get_plastic_strain(s2d_handle h, double* ep);

/* Visco-elasticity (kind='VISCO': generalized Maxwell body, MAT_VISCO_stress mat_visco.f90:206-248, through the same
 * strain -> stress -> force branch of MAT_Fint, mat_gen.f90:451-457).  Per material k: nbody(k) <= 8 mechanisms,
 * moduli(2,k) = lambda_inf, mu_inf (the unrelaxed moduli), wbody(8,k) relaxation frequencies, theta(8,3,k) as
 * get_attenuation (mat_visco.f90:251-340) returns them -- the host evaluates that routine (a least-squares fit per
 * material), the device keeps the memory variables el(ngll,ngll,Nbody,3) and the previous strain of every element and
 * advances them in every force evaluation.  elem_set(nelem), natural element order: 0 = elastic element.
 * s2d_cart_set_material still gives rho, cp, cs (mass, absorbing boundaries, Courant step use the input speeds, as the
 * reference does).  P-SV, ngll <= 6, no Kelvin-Voigt or plastic elements in the same problem. */
/* Damage rheology (kind='DMG': Lyakhovsky et al. 1997 / Hamiel et al. 2004 as in mat_damage.f90; MAT_DMG_stress :337-445,
 * compute_stress :453-491, MAT_DMG_init_elem_work :209-279).  par(13,nsets) = lambda, mu (intact), phi [degrees], alpha
 * (initial damage), Cd, beta, R, e0(3), ep(3) of every DMG material; elem_set(nelem), natural element order, 0 = elastic
 * element.  The device keeps alpha and the plastic strain ep of every element GLL point (s2d_cart_get_damage_state:
 * (ngll,ngll,4,nelem) = alpha, ep11, ep22, ep12) and advances them in every force evaluation; when compute_stress's
 * loss-of-convexity checks fail the run stops with "MAT_DMG: damage exceeded critical value", as the reference does.
 * P-SV, ngll <= 6, no Kelvin-Voigt, plastic or visco-elastic elements in the same problem. */
int s2d_cart_set_damage(s2d_handle h, int32_t nsets, const double* par, const int32_t* elem_set);
int s2d_cart_get_damage_state(s2d_handle h, double* state);
int s2d_cart_set_visco(s2d_handle h, int32_t nsets, const int32_t* nbody, const double* moduli, const double* wbody,
                       const double* theta, const int32_t* elem_set);

/* ---- device-side structured builder (mesh_cartesian.f90:219-314 + init on the GPU) --------- */
/* Builds, directly in HBM, a MESH_CART problem: nx*nz Q4 elements on [x0,x1]x[z0,z1], optional
 * horizontal split-node fault after element row ezflt (0 = none), natural (row-major) element
 * order, GLL numbering by the rule of SE_init_numbering (spec_grid.f90:198-314), flat-grid
 * coefficient planes (mat_elastic.f90:323-358) and mass (mat_mass.f90:50-57).
 * Material: seed != 0 -> the heterogeneous hash model of the synthetic benchmark (cs,cp,rho per
 * global GLL lattice point; lattice origin (ix0,iz0) so x-strips of one global mesh agree),
 * seed == 0 -> homogeneous rho,cp,cs.  Boundaries are added afterwards with the s2d_cart_add_*
 * calls, then s2d_commit. */
typedef struct {
  int32_t ngll, ndof, nx, nz, ezflt;
  double x0, x1, z0, z1;
  uint64_t seed;
  int64_t ix0, iz0;       /* lattice offset of this strip inside the global mesh */
  double rho, cp, cs;     /* homogeneous material when seed == 0 */
  int32_t precision;      /* 8 | 4 */
  s2d_scheme scheme;      /* dt <= 0: dt = courant / max(c/dx) as CHECK_grid + TIME_init do */
  double courant;
  int32_t device;
  int32_t halo_left, halo_right; /* strip has a neighbour on that side (no physical boundary there) */
  /* 0: isotropic P-SV boxes keep only (lambda, mu) per GLL point in HBM and the strip kernel forms
   *    the six planes of MAT_ELAST_init_a (mat_elastic.f90:334-340,355-357) in registers, with the
   *    same sequence of roundings (a third of the coefficient traffic);
   * 1: all nelast planes are stored, as matwrk_elast_type%a is (mat_elastic.f90:11-14). */
  int32_t coef_mode;
  /* != 0: OPT_RENUMBER = .true., the reference's default (SRC/constants.f90:10-15): the elements are put in reverse
   * Cuthill-McKee order (MESH_STRUCTURED_renumber, mesh_structured.f90:204-269; genrcm, rcm.f90) and the GLL nodes are
   * numbered by SE_init_numbering in that order, so that s2d_cart_get's ibool, s2d_get_fields / s2d_set_fields and
   * every node-ordered output are those of a stock reference build.  0: natural (row by row) element order.
   * s2d_cart_set_material / s2d_cart_set_kv_elems keep the natural element order either way.  Not on x-strips. */
  int32_t renumber;
} s2d_cart_desc;
int s2d_cart_create(s2d_handle* h, const s2d_cart_desc* desc);
/* BC_ABSO_init on mesh side tag 1 bottom, 2 right, 3 top, 4 left (mesh_structured.f90:86-196);
 * adds CoefA2Vrhs*C to the mass like bc_abso.f90:243.  Must precede s2d_cart_add_fault when the
 * reference deck lists ABSORB first. */
int s2d_cart_add_abso(s2d_handle h, int32_t side_tag, int32_t stacey);
/* BC_PERIO_init between two opposite sides of the box: tags (4,2) / (2,4) or (1,3) / (3,1); sums the
 * mass of the pairs (bc_periodic.f90:73).  Must precede the other boundaries, as in bc_gen.f90:221-228.
 * Not available across x-strips (a strip with a neighbour has no left / right physical side). */
int s2d_cart_add_periodic(s2d_handle h, int32_t master_tag, int32_t slave_tag);
/* two-sided DYNFLT on tags 5,6 with linear slip weakening: uniform Dc, MuS, MuD, Tn, Tt, and Tt_nuc
 * where |x - x_nuc| <= half_nuc (the PWCONR patch of the TPV3 deck). */
int s2d_cart_add_fault_swf(s2d_handle h, double Dc, double MuS, double MuD, double Tn, double Tt,
                           double Tt_nuc, double x_nuc, double half_nuc, int32_t oixd, int32_t oitd,
                           int32_t nt_max, int32_t* fault_id);
int s2d_cart_add_force(s2d_handle h, double x, double z, const double dir[2], int32_t* src_id);
/* SRC_MOMENT_init on the box: moment tensor M(2,ndof) column-major as so%M (src_moment.f90:44-100) at the
 * GLL node nearest to (x,z) */
int s2d_cart_add_moment(s2d_handle h, double x, double z, const double* M, int32_t* src_id);
int s2d_cart_add_receivers(s2d_handle h, int32_t nx, double xa, double za, double xb, double zb,
                           char field, int32_t isamp, int32_t nt_rec);
/* General BC_DYNFLT_init on the box (bc_dynflt.f90:231-520).  tags (5,6): the split-node row of ezflt;
 * (1,0) / (3,0): the bottom / top side as a one-sided fault (symmetry assumption, tags(2)=0, :700-716).
 * s2d_cart_fault_nodes gives np and bc%coord(2,np) so that the host can evaluate its distributions there;
 * `law` then carries np, T0(np,2), cohesion, V0, the friction-law arrays (swf_*, rsf_*, twf_*), the
 * normal-stress law and the output strides.  Its topology members (node1, node2, n1, B, invM1, invM2, Z,
 * coord, CoefA2V, CoefA2D) are ignored: the builder fills them. */
int s2d_cart_fault_nodes(s2d_handle h, int32_t tag1, int32_t tag2, int32_t* np, double* coord);
int s2d_cart_add_dynflt(s2d_handle h, int32_t tag1, int32_t tag2, const s2d_dynflt_desc* law, int32_t* fault_id);
/* bc_DIRNEU_init on side tag 1..4: kind 1 = Neumann (homogeneous), 2 = Dirichlet, per component */
int s2d_cart_add_dirneu(s2d_handle h, int32_t side_tag, int32_t kind_h, int32_t kind_v);
/* a fault as BC_DYNFLT_init left it (bc_dynflt.f90:392-458): node count, bc%coord(2,np), bc%T0(np,2),
 * bc%B(np,1) -- what FltXX_sem2d.hdr and FltXX_init_sem2d.tab hold.  Any pointer may be NULL. */
int s2d_cart_fault_info(s2d_handle h, int32_t fault_id, int32_t* np, double* coord, double* T0, double* B);
/* number of stations kept (duplicates dropped, receivers.f90:255) and their relocated positions
 * rec%coord(2,nx) (receivers.f90:231-303); coord may be NULL */
int s2d_cart_receiver_info(s2d_handle h, int32_t* nx, double* coord);
/* REC_LINE with AtNode=F: stations stay where they are, sampled with the Lagrange interpolant of the element
 * their nearest node picks (receivers.f90:262-300) */
int s2d_cart_add_receivers_interp(s2d_handle h, int32_t nx, double xa, double za, double xb, double zb,
                                  char field, int32_t isamp, int32_t nt_rec);
/* Kelvin-Voigt viscosity of the whole box (mat_kelvin_voigt.f90:117-150): eta(npoin) at the GLL nodes in the
 * caller's numbering, already multiplied by dt when ETAxDT.  eta must be a function of position (a constant
 * or a distribution evaluated at the node coordinates), which makes the element-wise d + eta*v node-wise.
 * With it the node update runs in separate passes (no fused step). */
int s2d_cart_set_kv(s2d_handle h, const double* eta_node);
/* Caller-supplied material on the structured builder: what MAT_getProp returns once MAT_read / MAT_init_prop have
 * run (SRC/mat_gen.f90:101-303) -- every tag with its own material, every property a constant or a DIST_* field --
 * evaluated by the host at the GLL points: rho, cp, cs (ngll,ngll,nelem), elements in the box's natural order
 * (element (ix,iz), 0-based, at ix + nx*iz; the reference's order before MESH_STRUCTURED_renumber).  Rebuilds the
 * coefficient planes (MAT_ELAST_init_a, mat_elastic.f90:290-360), the mass (mat_mass.f90:29-61) and, unless the
 * scheme came with a dt, the time step (init.f90:187-225, time.f90:334-341).  Before any boundary condition. */
int s2d_cart_set_material(s2d_handle h, const double* rho, const double* cp, const double* cs);
/* matwrk_kv_type%eta (SRC/mat_kelvin_voigt.f90:117-150) of the Kelvin-Voigt elements of the box: eta(ngll,ngll,nkv),
 * already times dt when ETAxDT, elem_ids(nkv) 1-based in natural element order.  The force kernel then sees
 * d + eta*v element by element (:137-150).  Replaces a previous s2d_cart_set_kv. */
int s2d_cart_set_kv_elems(s2d_handle h, int32_t nkv, const int32_t* elem_ids, const double* eta);
/* &GENERAL W (SRC/input.f90:82,119): finite seismogenic width of a 2.5D run.  The builder forms
 * beta = dvol * mu * (pi (1 - nu) / W)^2, nu = lambda / (lambda + mu) / 2 (P-SV; SH without the (1 - nu) factor)
 * at every GLL point (MAT_ELAST_init_25D, mat_elastic.f90:363-383) from the current material.  Before commit. */
int s2d_cart_set_w25d(s2d_handle h, double W);
int s2d_cart_info(s2d_handle h, int64_t* npoin, int64_t* nelem, double* dt);
/* Overrides the time step (time%dt) before any boundary is added.  x-strips of one global mesh
 * must agree on dt: the host takes the minimum of the per-strip Courant steps (the reference takes
 * the maximum of c/dx over the whole grid, init.f90:187-225) and sets it on every strip. */
int s2d_cart_set_dt(s2d_handle h, double dt);
/* Seeded non-trivial state for measurements and windowed parity checks (no reference counterpart: the
 * reference starts from rest or from a restart file): displ = amp_d*u, veloc = amp_v*u', accel = 0, with
 * u, u' in U(-1,1) from the same counter-based hash as the synthetic medium, keyed by the node's global
 * lattice coordinates (partition-independent; the two sides of a split fault row get equal values). */
int s2d_cart_fill_fields(s2d_handle h, uint64_t seed, double amp_d, double amp_v);
/* A rectangular window of the GLL lattice (columns gx0..gx0+nwx-1, rows gz0..gz0+nwz-1; a split fault row is
 * two lattice rows, lower side first): out[c][row][col] in FP64.  What fields%displ/veloc/accel hold there,
 * without moving the whole (npoin,ndof) arrays of s2d_get_fields.  Any pointer may be NULL. */
int s2d_cart_get_window(s2d_handle h, int32_t gx0, int32_t gz0, int32_t nwx, int32_t nwz, double* displ,
                        double* veloc, double* accel);
/* get_GLL_info (SRC/gll.f90:19-36) as the builder computed it: xgll(ngll), wgll(ngll), hprime(ngll,ngll) column-major
 * with hprime(i,j) = h'_i(x_j).  Any pointer may be NULL. */
int s2d_cart_get_gll(s2d_handle h, double* xgll, double* wgll, double* hprime);
/* PLOT_FIELD's element-wise snapshot fields (SRC/plot_gen.f90:239-300), computed on the device from the resident
 * fields: what = 'E' strain (e11, e22, e12 | e13, e23; FIELD_strain_elem, fields.f90:192-237), 'S' stress (s11, s22,
 * s12 | s13, s23; MAT_stress_dv, mat_gen.f90:626-641, with the Kelvin-Voigt d + eta*v; isotropic MAT_ELAST_stress,
 * mat_elastic.f90:822-839), 'd' divergence and 'c' curl of the velocity (P-SV only; fields.f90:242-285).
 * out[ncomp][nelem][ngll*ngll] float32 -- one record of ngll*ngll reals per element and component, the layout of
 * e11_NNN_sem2d.dat etc. -- elements in the caller's order (natural, or RCM with `renumber`). */
int s2d_cart_snapshot_elem(s2d_handle h, char what, float* out);
/* copies of builder outputs for parity tests against the oracle (host pointers, may be NULL) */
int s2d_cart_get(s2d_handle h, int32_t* ibool, double* a, double* rmass, double* coord);

/* ---- x-strip halo (multi-GPU) --------------------------------------------------- */
/* Number of values in one interface column (nodes * ndof), and device pointers of the send /
 * receive staging buffers: [0] left, [1] right.  After every force evaluation the engine packs
 * its partial sums of the interface nodes into send[side], calls the registered exchange hook,
 * and adds recv[side] (left partial + right partial on both ranks, so both stay bit-identical). */
int s2d_halo_info(s2d_handle h, int64_t* count, void** send_dev, void** recv_dev);
typedef int (*s2d_exchange_fn)(void* user, void* stream);
int s2d_halo_set_exchange(s2d_handle h, s2d_exchange_fn fn, void* user);
/* Peer-memory variant (NVLink, no host hook and no NCCL call on the step path): the engine's pack
 * kernel writes its partial sums straight into the NEIGHBOUR's receive slot and then raises the
 * neighbour's flag with the sequence number of the force evaluation; the neighbour's stream waits on
 * its own flag before it adds them.  The pointers are device pointers valid on this GPU:
 *   left_recv_dev  = the left neighbour's recv[1],  left_flag_dev  = &(its flags)[1]
 *   right_recv_dev = the right neighbour's recv[0], right_flag_dev = &(its flags)[0]
 * as returned by s2d_halo_peer_buffers on the neighbour (same process, peer access enabled), or
 * mapped through CUDA IPC by s2d_halo_ipc_export / s2d_halo_ipc_open (one process per GPU).
 * Every strip must run the same sequence of force evaluations. */
int s2d_halo_set_peers(s2d_handle h, void* left_recv_dev, void* right_recv_dev,
                       void* left_flag_dev, void* right_flag_dev);
/* recv_dev[2]: receive buffers (2 slots of s2d_halo_info's count each) [0] left, [1] right;
 * flags_dev: two 64-bit flags, [0] raised by the left neighbour, [1] by the right one */
int s2d_halo_peer_buffers(s2d_handle h, void** recv_dev, void** flags_dev);
/* blob: S2D_HALO_IPC_BYTES bytes (three cudaIpcMemHandle_t) to hand to both neighbours */
#define S2D_HALO_IPC_BYTES 192
int s2d_halo_ipc_export(s2d_handle h, void* blob);
/* the neighbours' blobs (NULL where there is no neighbour); calls s2d_halo_set_peers */
int s2d_halo_ipc_open(s2d_handle h, const void* left_blob, const void* right_blob);

#ifdef __cplusplus
}
#endif
#endif /* SEM2D_B200_H */
