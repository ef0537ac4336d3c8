! sem2d_b200_exporters.f90 -- the routines a SEM2DPACK maintainer adds INSIDE the reference's modules
! so that their private components reach the C-ABI (INTEGRATION.md, sections 1 and 3).
!
! Delivered as source, not compiled here (no Fortran compiler in the build image).  Each block names
! the module it belongs to; paste it before that module's "end module".  `use sem2d_b200` (the
! interface module, fortran/sem2d_b200.f90) and `use iso_c_binding` are assumed in every block.
! The call order is the one of init_main (SRC/init.f90:16-131): MAT, then BC in input order with the
! periodic ones first (SRC/bc_gen.f90:221-246), then sources, then receivers.

!=======================================================================================
! module mat_gen  (SRC/mat_gen.f90): coefficient blocks elast%a and Kelvin-Voigt eta
!---------------------------------------------------------------------------------------
  subroutine MAT_export_b200(h, matwrk, grid)
    use spec_grid, only : sem_grid_type
    use mat_elastic, only : MAT_ELAST_export_a     ! returns a pointer to matwrk_elast_type%a (private)
    use mat_kelvin_voigt, only : MAT_KV_export_eta ! returns a pointer to matwrk_kv_type%eta (private)
    use constants, only : OPT_NGLL
    type(c_ptr), intent(in) :: h
    type(matwrk_elem_type), intent(in), target :: matwrk(:)
    type(sem_grid_type), intent(in) :: grid
    double precision, pointer :: a(:,:,:), eta(:,:)
    double precision, allocatable, target :: abuf(:,:,:,:), etabuf(:,:,:), betabuf(:,:,:)
    integer(c_int32_t), allocatable, target :: elem2set(:), kv_elem(:)
    integer :: e, nelast, nsets, nkv, ngll

    ngll = grid%ngll
    ! homogeneous elements share one work structure (mat_gen.f90:357-365): one coefficient set for them
    nsets = 0
    allocate(elem2set(grid%nelem))
    do e = 1, grid%nelem
      if (e > 1 .and. associated(matwrk(e)%elast, matwrk(1)%elast)) then
        elem2set(e) = elem2set(1)
      else
        nsets = nsets + 1
        elem2set(e) = nsets          ! 1-based block index
      endif
    enddo
    call MAT_ELAST_export_a(matwrk(1)%elast, a)
    nelast = size(a, 3)              ! 2/3 SH, 6/10 P-SV (mat_elastic.f90:255-268)
    allocate(abuf(ngll, ngll, nelast, nsets))
    do e = 1, grid%nelem
      call MAT_ELAST_export_a(matwrk(e)%elast, a)
      abuf(:,:,:,elem2set(e)) = a
    enddo
    if (grid%W < huge(1d0)) then     ! 2.5D: matwrk%elast%beta (mat_elastic.f90:280-284), one block per coefficient set
      allocate(betabuf(ngll, ngll, nsets))
      do e = 1, grid%nelem
        call MAT_ELAST_get_beta(matwrk(e)%elast, betabuf(:,:,elem2set(e)))
      enddo
      call s2d_check(h, s2d_set_elastic(h, nelast, nsets, abuf, elem2set, c_loc(betabuf), merge(1, 0, ngll == OPT_NGLL)))
    else
      call s2d_check(h, s2d_set_elastic(h, nelast, nsets, abuf, elem2set, c_null_ptr, merge(1, 0, ngll == OPT_NGLL)))
    endif

    nkv = count( (/ (associated(matwrk(e)%kv), e = 1, grid%nelem) /) )
    if (nkv > 0) then
      allocate(kv_elem(nkv), etabuf(ngll, ngll, nkv))
      nkv = 0
      do e = 1, grid%nelem
        if (.not. associated(matwrk(e)%kv)) cycle
        nkv = nkv + 1
        kv_elem(nkv) = e
        call MAT_KV_export_eta(matwrk(e)%kv, eta)   ! already multiplied by dt (mat_kelvin_voigt.f90:127)
        etabuf(:,:,nkv) = eta
      enddo
      call s2d_check(h, s2d_set_kv(h, nkv, kv_elem, etabuf))
    endif
  end subroutine MAT_export_b200

!=======================================================================================
! module bc_gen  (SRC/bc_gen.f90): dispatch in the order of BC_init
!---------------------------------------------------------------------------------------
  subroutine BC_export_b200(h, bc)
    type(c_ptr), intent(in) :: h
    type(bc_type), pointer :: bc(:)
    integer :: i
    if (.not. associated(bc)) return
    do i = 1, size(bc)                       ! periodic boundaries first (bc_gen.f90:221-228)
      if (bc(i)%kind == IS_PERIOD) call BC_PERIO_export_b200(h, bc(i)%perio)
    enddo
    do i = 1, size(bc)                       ! then input order (bc_gen.f90:229-246)
      select case (bc(i)%kind)
        case (IS_ABSORB); call BC_ABSO_export_b200(h, bc(i)%abso)
        case (IS_DIRNEU); call BC_DIRNEU_export_b200(h, bc(i)%dirneu)
        case (IS_DYNFLT); call BC_DYNFLT_export_b200(h, bc(i)%dynflt)
        case (IS_PERIOD, IS_EMPTY)
        case default
          call IO_abort('BC_export_b200: this boundary kind is not on the B200 path')
      end select
    enddo
  end subroutine BC_export_b200

!=======================================================================================
! module bc_periodic  (SRC/bc_periodic.f90)
!---------------------------------------------------------------------------------------
  subroutine BC_PERIO_export_b200(h, bc)
    type(c_ptr), intent(in) :: h
    type(bc_periodic_type), intent(in) :: bc
    call s2d_check(h, s2d_add_periodic(h, bc%master%npoin, bc%master%node, bc%slave%node))
  end subroutine BC_PERIO_export_b200

!=======================================================================================
! module bc_abso  (SRC/bc_abso.f90): bc_abso_type has private components (:38-46)
!---------------------------------------------------------------------------------------
  subroutine BC_ABSO_export_b200(h, bc)
    type(c_ptr), intent(in) :: h
    type(bc_abso_type), intent(in) :: bc
    double precision :: dummy(1)
    integer(c_int32_t) :: idummy(1)
    if (bc%stacey) then
      if (bc%periodic) call IO_abort('BC_ABSO_export_b200: Stacey + periodic is not on the B200 path')
      call s2d_check(h, s2d_add_abso(h, bc%topo%npoin, bc%topo%node, bc%C, merge(1,0,bc%is_flat), bc%n, 1, &
                                     bc%topo%nelem, bc%topo%ibool, bc%K))
    else
      call s2d_check(h, s2d_add_abso(h, bc%topo%npoin, bc%topo%node, bc%C, merge(1,0,bc%is_flat), bc%n, 0, &
                                     0, idummy, dummy))
    endif
  end subroutine BC_ABSO_export_b200

!=======================================================================================
! module bc_dirneu  (SRC/bc_dirneu.f90): kinds 1 = Neumann, 2 = Dirichlet per component (:17-23)
!---------------------------------------------------------------------------------------
  subroutine BC_DIRNEU_export_b200(h, bc)
    type(c_ptr), intent(in) :: h
    type(bc_dirneu_type), intent(in), target :: bc
    type(c_ptr) :: bh, bv
    bh = c_null_ptr
    bv = c_null_ptr
    if (associated(bc%hstf)) bh = c_loc(bc%B)     ! time-dependent Neumann: the host passes stf(t) per step
    if (associated(bc%vstf)) bv = c_loc(bc%B)     ! in s2d_step's bc_ampli table
    call s2d_check(h, s2d_add_dirneu(h, bc%topo%npoin, bc%topo%node, bc%kind(1), bc%kind(2), bh, bv))
  end subroutine BC_DIRNEU_export_b200

!=======================================================================================
! module bc_dynflt  (SRC/bc_dynflt.f90): bc_dynflt_type (:18-38) after BC_DYNFLT_init (:231-520)
!---------------------------------------------------------------------------------------
  subroutine BC_DYNFLT_export_b200(h, bc)
    type(c_ptr), intent(in) :: h
    type(bc_dynflt_type), intent(inout), target :: bc
    type(s2d_dynflt_desc), target :: d            ! bind(C) mirror of the struct in include/sem2d_b200.h
    integer(c_int32_t) :: fault_id
    d%np = bc%npoin
    d%node1 = c_loc(bc%node1)
    d%node2 = c_null_ptr                          ! one-sided (symmetric) fault: bc2 not associated (:700-716)
    if (associated(bc%bc2)) d%node2 = c_loc(bc%node2)
    d%n1 = c_loc(bc%n1);  d%B = c_loc(bc%B);  d%invM1 = c_loc(bc%invM1)
    d%invM2 = c_null_ptr
    if (associated(bc%bc2)) d%invM2 = c_loc(bc%invM2)
    d%Z = c_loc(bc%Z);  d%T0 = c_loc(bc%T0);  d%cohesion = c_loc(bc%cohesion);  d%coord = c_loc(bc%coord)
    d%V0 = c_loc(bc%V)
    d%CoefA2V = bc%CoefA2V;  d%CoefA2D = bc%CoefA2D
    d%allow_opening = merge(1, 0, bc%allow_opening)
    call swf_export_b200(bc%swf, d)               ! each law's module fills its own members the same way
    call rsf_export_b200(bc%rsf, d)
    call twf_export_b200(bc%twf, d)
    call normal_export_b200(bc%normal, d)
    d%oix1 = bc%oix1;  d%oixn = bc%oixn;  d%oixd = bc%oixd;  d%oit = bc%oit;  d%oitd = bc%oitd
    d%nt_max = bc%nt                              ! time%nt, kept by BC_DYNFLT_init for NSAMP (:481)
    call s2d_check(h, s2d_add_dynflt(h, c_loc(d), fault_id))
    bc%gpu_id = fault_id                          ! used by BC_DYNFLT_write to fetch the records (s2d_get_fault)
  end subroutine BC_DYNFLT_export_b200

!=======================================================================================
! module src_gen  (SRC/src_gen.f90) with src_force / src_moment exporters
!---------------------------------------------------------------------------------------
  subroutine SO_export_b200(h, so)
    type(c_ptr), intent(in) :: h
    type(source_type), pointer :: so(:)
    integer :: k
    if (.not. associated(so)) return
    do k = 1, size(so)
      select case (so(k)%mech%kind)
        case (tag_force);  call FORCE_export_b200(h, so(k)%mech%force, so(k)%gpu_id)
        case (tag_moment); call SRC_MOMENT_export_b200(h, so(k)%mech%moment, so(k)%gpu_id)
        case default
          call IO_abort('SO_export_b200: incident-wave sources are not on the B200 path')
      end select
    enddo
  end subroutine SO_export_b200
  ! module src_force:   call s2d_check(h, s2d_add_force(h, so%iglob, so%dir, id))
  ! module src_moment:  the terms of SRC_MOMENT_add (:183-197) flattened in its own loop order
  subroutine SRC_MOMENT_export_b200(h, so, id)
    type(c_ptr), intent(in) :: h
    type(so_moment_type), intent(in) :: so
    integer(c_int32_t), intent(out) :: id
    integer(c_int32_t), allocatable :: node(:)
    double precision, allocatable :: coef(:,:)
    integer :: k, nel, ngll, ndof, t
    ngll = size(so%iglob_xi, 1);  nel = size(so%iglob_xi, 2);  ndof = size(so%coef_xi, 2)
    allocate(node(2*ngll*nel), coef(2*ngll*nel, ndof))
    t = 0
    do k = 1, nel
      node(t+1:t+ngll) = so%iglob_xi(:,k);   coef(t+1:t+ngll,:) = so%coef_xi(:,:,k);   t = t + ngll
      node(t+1:t+ngll) = so%iglob_eta(:,k);  coef(t+1:t+ngll,:) = so%coef_eta(:,:,k);  t = t + ngll
    enddo
    call s2d_check(h, s2d_add_moment(h, t, node, coef, id))
  end subroutine SRC_MOMENT_export_b200

!=======================================================================================
! module receivers  (SRC/receivers.f90): rec_type is private (:9-20)
!---------------------------------------------------------------------------------------
  subroutine REC_export_b200(h, rec)
    type(c_ptr), intent(in) :: h
    type(rec_type), intent(in), target :: rec
    if (rec%AtNode) then
      call s2d_check(h, s2d_add_receivers(h, rec%nx, rec%SeisField, rec%isamp, rec%nt, 1, &
                                          c_loc(rec%iglob), c_null_ptr, c_null_ptr))
    else
      call s2d_check(h, s2d_add_receivers(h, rec%nx, rec%SeisField, rec%isamp, rec%nt, 0, &
                                          c_null_ptr, c_loc(rec%einterp), c_loc(rec%interp)))
    endif
  end subroutine REC_export_b200
  ! REC_write (:351-392) is unchanged: before it, main.f90 calls s2d_get_seis(h, rec%sis).

!=======================================================================================
! module solver  (SRC/solver.f90): solve(pb) on the B200 path
!---------------------------------------------------------------------------------------
  subroutine solve_b200(pb, nsteps)
    ! nsteps passes of the loop body of main.f90:51-99 (solve + REC_store + BC_write) on the device.
    ! The source time functions are evaluated here exactly where SO_add evaluates them
    ! (src_gen.f90:300-303); HHT-alpha uses t_alpha, symplectic schemes one row per stage.
    type(problem_type), intent(inout) :: pb
    integer, intent(in) :: nsteps
    double precision, allocatable, target :: ampli(:,:)
    double precision :: t
    integer :: k, s, q, nsrc, nst
    nsrc = 0
    if (associated(pb%src)) nsrc = size(pb%src)
    nst = max(1, pb%time%nstages)
    allocate(ampli(max(nsrc,1), nsteps*nst))
    do k = 1, nsteps
      t = (pb%it + k) * pb%time%dt
      if (pb%time%nstages > 0) then
        t = t - pb%time%dt
        do q = 1, nst
          t = t + pb%time%dt * pb%time%a(q)
          do s = 1, nsrc
            ampli(s, (k-1)*nst + q) = SO_ampli(pb%src(s), t)     ! stf(t - tdelay) * ampli
          enddo
        enddo
      else
        if (pb%time%kind == 'HHT-alpha') t = t + (pb%time%alpha - 1d0) * pb%time%dt
        do s = 1, nsrc
          ampli(s, k) = SO_ampli(pb%src(s), t)
        enddo
      endif
    enddo
    call s2d_check(pb%gpu, s2d_step(pb%gpu, nsteps, c_loc(ampli), c_null_ptr))
    pb%it = pb%it + nsteps
    pb%time%time = pb%it * pb%time%dt
  end subroutine solve_b200
