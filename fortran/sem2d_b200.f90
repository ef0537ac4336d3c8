! sem2d_b200.f90 -- ISO_C_BINDING interface to libsem2d_b200.so (include/sem2d_b200.h).
!
! This is the module a SEM2DPACK maintainer adds to SRC/ (INTEGRATION.md has the call sites).  It is
! delivered as source: no Fortran compiler exists in the build image (SURVEY.md header), so it is not
! compiled here; the same C entry points are exercised by the C++ host (host/) and by ctypes
! (sem2dpack_b200/capi.py) in every GPU test.  Arrays are passed as they are in the reference:
! column-major, 1-based node / element ids, default integers, double precision.
module sem2d_b200
  use iso_c_binding
  implicit none
  type, bind(C) :: s2d_scheme            ! timescheme_type, time.f90:5-11
    integer(c_int32_t) :: kind           ! 0 leapfrog (solver.f90:140-160), 1 newmark (:42-84),
                                         ! 2 HHT-alpha (:89-128), 3 symplectic (:169-199)
    real(c_double) :: dt, beta, gamma, alpha
    integer(c_int32_t) :: nstages        ! time%nstages, time%a(1:nstages+1), time%b(1:nstages)
    real(c_double) :: coa(9), cob(8)
  end type
  type, bind(C) :: s2d_cart_desc          ! include/sem2d_b200.h, same member order
    integer(c_int32_t) :: ngll, ndof, nx, nz, ezflt
    real(c_double) :: x0, x1, z0, z1
    integer(c_int64_t) :: seed            ! 0: homogeneous rho, cp, cs below (then s2d_cart_set_material for anything else)
    integer(c_int64_t) :: ix0, iz0
    real(c_double) :: rho, cp, cs
    integer(c_int32_t) :: precision
    type(s2d_scheme) :: scheme
    real(c_double) :: courant
    integer(c_int32_t) :: device
    integer(c_int32_t) :: halo_left, halo_right
    integer(c_int32_t) :: coef_mode
    integer(c_int32_t) :: renumber        ! 1 = OPT_RENUMBER (constants.f90:10-15): node ids of a stock reference build
  end type
  interface
    integer(c_int) function s2d_create(h, ngll, ndof, nelem, npoin, ibool, hprime, rmass, precision, scheme, device) &
        bind(C, name='s2d_create')
      import
      type(c_ptr), intent(out) :: h
      integer(c_int32_t), value :: ngll, ndof, nelem, npoin, precision, device
      integer(c_int32_t), intent(in) :: ibool(*)        ! grid%ibool(ngll,ngll,nelem)   spec_grid.f90:52
      real(c_double), intent(in) :: hprime(*), rmass(*) ! grid%hprime, pb%rmass(npoin,ndof)
      type(s2d_scheme), intent(in) :: scheme
    end function
    integer(c_int) function s2d_set_elastic(h, nelast, ncoefsets, a, elem2set, beta25d, kd2) bind(C, name='s2d_set_elastic')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: nelast, ncoefsets, kd2
      real(c_double), intent(in) :: a(*)                ! a(ngll,ngll,nelast) blocks   mat_elastic.f90:290-360
      integer(c_int32_t), intent(in) :: elem2set(*)     ! 1-based block of each element
      type(c_ptr), value :: beta25d                     ! c_loc(beta(ngll,ngll,ncoefsets)) when grid%W is finite
                                                        ! (mat_elastic.f90:280-284,363-383), else c_null_ptr
    end function
    integer(c_int) function s2d_kernel_route(h, route) bind(C, name='s2d_kernel_route')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), intent(out) :: route          ! 1: the box was recognised in ibool, strip kernel; 0: any-mesh kernel
    end function
    integer(c_int) function s2d_set_kv(h, nkv, elem_ids, eta) bind(C, name='s2d_set_kv')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: nkv
      integer(c_int32_t), intent(in) :: elem_ids(*)
      real(c_double), intent(in) :: eta(*)              ! eta(ngll,ngll) per KV element, already *dt
    end function
    integer(c_int) function s2d_add_abso(h, np, node, C, is_flat, n, stacey, nbe, bibool, K) bind(C, name='s2d_add_abso')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: np, is_flat, stacey, nbe
      integer(c_int32_t), intent(in) :: node(*), bibool(*)
      real(c_double), intent(in) :: C(*), n(*), K(*)    ! bc_abso_type, bc_abso.f90:38-46
    end function
    integer(c_int) function s2d_add_periodic(h, np, master, slave) bind(C, name='s2d_add_periodic')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: np
      integer(c_int32_t), intent(in) :: master(*), slave(*)   ! bc%master%node, bc%slave%node   bc_periodic.f90:11-14
    end function
    integer(c_int) function s2d_add_dirneu(h, np, node, kind_h, kind_v, B_h, B_v) bind(C, name='s2d_add_dirneu')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: np, kind_h, kind_v
      integer(c_int32_t), intent(in) :: node(*)
      type(c_ptr), value :: B_h, B_v                     ! c_null_ptr = homogeneous Neumann
    end function
    integer(c_int) function s2d_add_dynflt(h, desc, fault_id) bind(C, name='s2d_add_dynflt')
      import
      type(c_ptr), value :: h
      type(c_ptr), value :: desc                         ! type(s2d_dynflt_desc), see the header
      integer(c_int32_t), intent(out) :: fault_id
    end function
    integer(c_int) function s2d_add_force(h, iglob, dir, src_id) bind(C, name='s2d_add_force')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: iglob
      real(c_double), intent(in) :: dir(2)
      integer(c_int32_t), intent(out) :: src_id
    end function
    integer(c_int) function s2d_add_moment(h, nterms, node, coef, src_id) bind(C, name='s2d_add_moment')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: nterms
      integer(c_int32_t), intent(in) :: node(*)          ! iglob_xi(:,k), iglob_eta(:,k), k = 1..nel   src_moment.f90:150-152
      real(c_double), intent(in) :: coef(*)              ! coef_xi(:,:,k), coef_eta(:,:,k) as (nterms,ndof)
      integer(c_int32_t), intent(out) :: src_id
    end function
    integer(c_int) function s2d_add_receivers(h, nx, field, isamp, nt_rec, at_node, iglob, einterp, interp) &
        bind(C, name='s2d_add_receivers')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: nx, isamp, nt_rec, at_node
      character(kind=c_char), value :: field
      type(c_ptr), value :: iglob, einterp, interp
    end function
    integer(c_int) function s2d_commit(h, variant) bind(C, name='s2d_commit')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: variant               ! 0 patch (default), 1 colour, 2 atomic
    end function
    integer(c_int) function s2d_set_fields(h, d, v, a) bind(C, name='s2d_set_fields')
      import
      type(c_ptr), value :: h
      real(c_double), intent(in) :: d(*), v(*), a(*)
    end function
    integer(c_int) function s2d_get_fields(h, d, v, a) bind(C, name='s2d_get_fields')
      import
      type(c_ptr), value :: h
      real(c_double), intent(out) :: d(*), v(*), a(*)
    end function
    integer(c_int) function s2d_step(h, nsteps, src_ampli, bc_ampli) bind(C, name='s2d_step')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: nsteps
      type(c_ptr), value :: src_ampli, bc_ampli          ! (nsrc,nsteps), (2*ndirneu,nsteps) or c_null_ptr
    end function
    integer(c_int) function s2d_get_seis(h, sis) bind(C, name='s2d_get_seis')
      import
      type(c_ptr), value :: h
      real(c_float), intent(out) :: sis(*)               ! rec%sis(nt,nx,ndof)   receivers.f90:12
    end function
    integer(c_int) function s2d_get_fault(h, fault_id, records, nout, potency, ncalls) bind(C, name='s2d_get_fault')
      import
      type(c_ptr), value :: h
      integer(c_int32_t), value :: fault_id
      type(c_ptr), value :: records, potency
      integer(c_int32_t), intent(out) :: nout, ncalls
    end function
    integer(c_int) function s2d_progress(h, vmax, dmax) bind(C, name='s2d_progress')
      import
      type(c_ptr), value :: h
      real(c_double), intent(out) :: vmax, dmax
    end function
    integer(c_int) function s2d_destroy(h) bind(C, name='s2d_destroy')
      import
      type(c_ptr), value :: h
    end function
    type(c_ptr) function s2d_last_error(h) bind(C, name='s2d_last_error')
      import
      type(c_ptr), value :: h
    end function
    ! ---- structured builder: a MESH_CART problem made in HBM from the &MESH_CART / &MATERIAL values -----------------
    ! (what host/sem2d_host.hpp:init_main does in C++; element-wise arrays in natural element order)
    integer(c_int) function s2d_cart_create(h, desc) bind(C, name='s2d_cart_create')
      import
      type(c_ptr), intent(out) :: h
      type(s2d_cart_desc), intent(in) :: desc
    end function
    integer(c_int) function s2d_cart_set_material(h, rho, cp, cs) bind(C, name='s2d_cart_set_material')
      import
      type(c_ptr), value :: h
      real(c_double), intent(in) :: rho(*), cp(*), cs(*)     ! (ngll,ngll,nelem): MAT_getProp at the GLL points
    end function
    integer(c_int) function s2d_cart_set_kv_elems(h, nkv, elem_ids, eta) bind(C, name='s2d_cart_set_kv_elems')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: nkv
      integer(c_int), intent(in) :: elem_ids(*)              ! 1-based
      real(c_double), intent(in) :: eta(*)                   ! matwrk%kv%eta(ngll,ngll) of those elements
    end function
    integer(c_int) function s2d_cart_set_w25d(h, W) bind(C, name='s2d_cart_set_w25d')
      import
      type(c_ptr), value :: h
      real(c_double), value :: W                             ! grid%W
    end function
    integer(c_int) function s2d_cart_set_plastic(h, nsets, par, elem_set) bind(C, name='s2d_cart_set_plastic')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: nsets
      real(c_double), intent(in) :: par(6,*)                 ! coh, phi, Tv, e0(3) of every PLAST material (mat_plastic.f90:66-118)
      integer(c_int), intent(in) :: elem_set(*)              ! 0 = elastic element, k = plastic material k
    end function
    integer(c_int) function s2d_cart_set_damage(h, nsets, par, elem_set) bind(C, name='s2d_cart_set_damage')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: nsets
      real(c_double), intent(in) :: par(13,*)                ! lambda, mu, phi, alpha, Cd, beta, R, e0(3), ep(3) (mat_damage.f90:108-173)
      integer(c_int), intent(in) :: elem_set(*)              ! 0 = elastic element, k = damage material k
    end function
    integer(c_int) function s2d_cart_get_damage_state(h, state) bind(C, name='s2d_cart_get_damage_state')
      import
      type(c_ptr), value :: h
      real(c_double), intent(out) :: state(*)                ! (ngll,ngll,4,nelem): matwrk%dmg%alpha, %ep(:,:,1:3)
    end function
    integer(c_int) function s2d_cart_set_visco(h, nsets, nbody, moduli, wbody, theta, elem_set) bind(C, name='s2d_cart_set_visco')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: nsets
      integer(c_int), intent(in) :: nbody(*)                 ! matwrk%visco%Nbody of every VISCO material
      real(c_double), intent(in) :: moduli(2,*)              ! lambda_inf, mu_inf (get_attenuation, mat_visco.f90:251-340)
      real(c_double), intent(in) :: wbody(8,*), theta(8,3,*) ! m%wbody(1:Nbody), m%theta(1:Nbody,1:3)
      integer(c_int), intent(in) :: elem_set(*)              ! 0 = elastic element, k = visco material k
    end function
    integer(c_int) function s2d_cart_get_plastic_strain(h, ep) bind(C, name='s2d_cart_get_plastic_strain')
      import
      type(c_ptr), value :: h
      real(c_double), intent(out) :: ep(*)                   ! matwrk%plast%ep(ngll,ngll,3) of every element (MAT_PLAST_export)
    end function
    integer(c_int) function s2d_cart_add_abso(h, side_tag, stacey) bind(C, name='s2d_cart_add_abso')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: side_tag, stacey
    end function
    integer(c_int) function s2d_cart_add_dirneu(h, side_tag, kind_h, kind_v) bind(C, name='s2d_cart_add_dirneu')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: side_tag, kind_h, kind_v
    end function
    integer(c_int) function s2d_cart_add_periodic(h, master_tag, slave_tag) bind(C, name='s2d_cart_add_periodic')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: master_tag, slave_tag
    end function
    integer(c_int) function s2d_cart_get(h, ibool, a, rmass, coord) bind(C, name='s2d_cart_get')
      import
      type(c_ptr), value :: h
      type(c_ptr), value :: ibool, a, rmass, coord           ! c_loc of the arrays wanted, c_null_ptr for the others
    end function
    integer(c_int) function s2d_cart_snapshot_elem(h, what, out) bind(C, name='s2d_cart_snapshot_elem')
      import
      type(c_ptr), value :: h
      character(kind=c_char), value :: what                  ! 'E' strain, 'S' stress, 'd' div, 'c' curl (plot_gen.f90:13)
      real(c_float), intent(out) :: out(*)
    end function
  end interface
contains
  subroutine s2d_check(h, rc)             ! error behaviour of the reference: IO_abort (stdio.f90:205-214)
    use stdio, only : IO_abort
    type(c_ptr), intent(in) :: h
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    character(len=256) :: text
    integer :: k
    if (rc == 0) return
    call c_f_pointer(s2d_last_error(h), msg, [256])
    text = ' '
    do k = 1, 256
      if (msg(k) == c_null_char) exit
      text(k:k) = msg(k)
    enddo
    call IO_abort(trim(text))
  end subroutine
end module sem2d_b200
