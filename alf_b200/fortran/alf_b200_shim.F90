!  alf_b200_shim.F90 -- the Fortran side of the drop-in boundary (SURVEY.md 8b): ALF's procedure NAMES and ARGUMENT LISTS, bodies forwarded to
!  libalf_b200.so through the generated ISO_C_BINDING interfaces of alf_b200_c_api.F90.  A maintainer compiles this file INSTEAD of the bodies of
!  Prog/wrapur_mod.F90, wrapul_mod.F90, cgr1_mod.F90, cgr2_2_mod.F90, QDRP_decompose_mod.F90, UDV_WRAP_mod.F90, Hop_mod.F90 (multiplications),
!  Wrapgr_mod.F90 (WRAPGRUP / WRAPGRDO) and tau_m_mod.F90; main.F90, every Hamiltonian, Operator_mod, Fields_mod, observables_mod stay untouched
!  (INTEGRATION.md shows the edits to Prog/Makefile).
!
!  NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler (SURVEY.md F2).  tests/test_fortran_shim_lint.py checks what can be
!  checked without one: the interface module is regenerated from include/alf_b200.h and must be identical, every C-ABI call below has the arity of
!  its prototype, every procedure that the reference exports is present with the reference's dummy-argument list, block constructs balance.
!
!  Two modes (SURVEY.md 8b):
!   (i)  compat mode  -- the procedures below, ONE chain (this MPI rank's) in a handle of n_chains = 1: every call ships its arrays to the device
!        and back.  Exact drop-in, meant for validation; the Metropolis decisions draw from the handle's own xoshiro256** stream seeded with the
!        rank's seed (ALF's RANDOM_NUMBER stream depends on the compiler runtime, SURVEY.md F5).
!   (ii) batched mode -- alf_b200_batched_sweep replaces the body of main.F90's get_sequential() branch (:714-887) for all chains of the handle.
module alf_b200_shim
  use iso_c_binding
  use alf_b200_c_api
  use runtime_error_mod          ! Terminate_on_error, ERROR_* codes (Libraries/Modules/runtime_error_mod.F90:56-130)
  use Operator_mod               ! type Operator (Prog/Operator_mod.F90:56-90)
  use Fields_mod                 ! type Fields   (Prog/Fields_mod.F90:79-99)
  use UDV_State_mod              ! type UDV_State (Prog/udv_state_mod.F90:85-110)
  use QMC_runtime_var           ! get_LOBS_ST, get_LOBS_EN (Prog/QMC_runtime_var_mod.F90)
  use Hamiltonian_main           ! Op_V, Op_T, nsigma, Ndim, N_FL, N_FL_eff, Calc_Fl_map, N_SUN, Ltrot, Symm, Projector, WF_L, WF_R (Hamiltonian_main_mod.F90:181-197)
  implicit none
  private
  public :: alf_b200_attach, alf_b200_detach, alf_b200_batched_sweep, alf_b200_reduce, alf_b200_handle_ptr
  public :: WRAPUR, WRAPUL, CGR, CGRP, CGR2_2, QDRP_decompose, UDV_Wrap_Pivot, decompose_UDV_state
  public :: Hop_mod_mmthr, Hop_mod_mmthr_m1, Hop_mod_mmthl, Hop_mod_mmthl_m1, Hop_mod_mmthlc, Hop_mod_Symm
  public :: WRAPGRUP, WRAPGRDO, TAU_M

  type(c_ptr), save :: alf_b200_handle_ptr = c_null_ptr
  integer(c_int), save :: alf_b200_device = 0      ! set by alf_b200_attach
  integer, parameter :: dp = kind(0.d0)

contains

  subroutine check(rc, file, line)        ! the C-ABI never aborts: map its return code to ALF's convention
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: file
    integer, intent(in) :: line
    if (rc /= 0) call Terminate_on_error(int(rc), file, line)
  end subroutine check

  !> Called once after ham%Ham_set (Prog/main.F90:318-319): flattens the PUBLIC fields of Op_V / Op_T (Prog/Operator_mod.F90:57-69) into the C tables.
  !> M_exp / E_exp / ExpOpT_vec are private in ALF and are rebuilt by alf_b200_finalize_model exactly as Op_set / Op_exp / Hop_mod_init do.
  !> n_chains = 1, seeds(1) = this rank's seed: compat mode; n_chains > 1: batched mode.
  subroutine alf_b200_attach(Nwrap, n_chains, device, seeds, Thtrot)
    integer, intent(in) :: Nwrap, n_chains, device
    integer, intent(in) :: seeds(:)
    integer, intent(in), optional :: Thtrot
    integer :: n, nf, stab
    alf_b200_device = device
    stab = 0
#if defined(STAB3)
    stab = 3
#endif
    call check(alf_b200_create(alf_b200_handle_ptr, Ndim, N_FL, N_SUN, Ltrot, Nwrap, size(Op_V,1), size(Op_T,1), merge(1,0,Symm), stab, n_chains, device), &
         & __FILE__, __LINE__)
    do nf = 1, N_FL
       do n = 1, size(Op_V,1)
          call check(alf_b200_set_op_v(alf_b200_handle_ptr, n, nf, Op_V(n,nf)%N, Op_V(n,nf)%N_non_zero, merge(1,0,Op_V(n,nf)%diag), Op_V(n,nf)%type, &
               & Op_V(n,nf)%P, Op_V(n,nf)%U, Op_V(n,nf)%E, dble(Op_V(n,nf)%g), aimag(Op_V(n,nf)%g), dble(Op_V(n,nf)%alpha), aimag(Op_V(n,nf)%alpha)), &
               & __FILE__, __LINE__)
          ! time-dependent coupling (Operator_mod.F90:66): the device then builds its vertex tables per time slice
          if (Op_V(n,nf)%get_g_t_alloc()) call check(alf_b200_set_op_v_gt(alf_b200_handle_ptr, n, nf, Op_V(n,nf)%g_t), __FILE__, __LINE__)
       enddo
       do n = 1, size(Op_T,1)
          call check(alf_b200_set_op_t(alf_b200_handle_ptr, n, nf, Op_T(n,nf)%N, merge(1,0,Op_T(n,nf)%diag), Op_T(n,nf)%P, Op_T(n,nf)%U, Op_T(n,nf)%E, &
               & dble(Op_T(n,nf)%g), aimag(Op_T(n,nf)%g)), __FILE__, __LINE__)
       enddo
    enddo
    if (Projector) then
       call check(alf_b200_set_projector(alf_b200_handle_ptr, Thtrot, size(WF_L(1)%P,2)), __FILE__, __LINE__)
       do nf = 1, N_FL
          call check(alf_b200_set_trial_wf(alf_b200_handle_ptr, nf, WF_L(nf)%P, WF_R(nf)%P), __FILE__, __LINE__)
       enddo
    endif
    call check(alf_b200_finalize_model(alf_b200_handle_ptr), __FILE__, __LINE__)
    call check(alf_b200_set_measure_interval(alf_b200_handle_ptr, get_LOBS_ST(), get_LOBS_EN()), __FILE__, __LINE__)      ! VAR_QMC, QMC_runtime_var_mod.F90:156-189
    call check(alf_b200_set_seeds(alf_b200_handle_ptr, int(seeds, c_int32_t)), __FILE__, __LINE__)
  end subroutine alf_b200_attach

  subroutine alf_b200_detach()
    integer(c_int) :: rc
    rc = alf_b200_destroy(alf_b200_handle_ptr)
    alf_b200_handle_ptr = c_null_ptr
  end subroutine alf_b200_detach

  !> Batched mode: the body of the get_sequential() branch, Prog/main.F90:714-887 (+ TAU_M when Ltau == 1), for all chains of the handle.
  !> nsigma_in(:,:,c) / nsigma_out(:,:,c) are chain c's nsigma%f before and after.
  subroutine alf_b200_batched_sweep(nsigma_in, nsigma_out, n_sweeps, ltau, obs, control)
    complex(kind=dp), intent(in)  :: nsigma_in(:,:,:)
    complex(kind=dp), intent(out) :: nsigma_out(:,:,:)
    integer, intent(in) :: n_sweeps, ltau
    real(kind=dp), intent(out) :: obs(:), control(16)
    call check(alf_b200_sweep_host(alf_b200_handle_ptr, n_sweeps, ltau, nsigma_in, nsigma_out, obs, control), __FILE__, __LINE__)
  end subroutine alf_b200_batched_sweep

  !> Replaces the MPI_REDUCE calls of Print_bin_Vec / Print_bin_Latt / Print_bin_Latt_Local (Prog/observables_mod.F90:425-438, 648-653, 828-834) and of
  !> Control_Print (Prog/control_mod.F90:397-452): NCCL over NVLink, in place on the device.  id(128) comes from alf_b200_comm_unique_id on rank 0
  !> and is distributed by the host program (MPI_BCAST in main.F90) before the first call with init = .true.
  subroutine alf_b200_reduce(nranks, rank, id, init, control)
    integer, intent(in) :: nranks, rank
    character(kind=c_char), intent(in) :: id(128)
    logical, intent(in) :: init
    real(kind=dp), intent(out) :: control(16)
    if (init) call check(alf_b200_comm_init(alf_b200_handle_ptr, nranks, rank, id), __FILE__, __LINE__)
    call check(alf_b200_reduce_bins(alf_b200_handle_ptr, 0), __FILE__, __LINE__)
    call check(alf_b200_reduce_control(alf_b200_handle_ptr, 0, control), __FILE__, __LINE__)
  end subroutine alf_b200_reduce

  ! ----------------------------------------------------------------------------------------------------------- compat mode helpers
  subroutine push_fields()               ! the device works on ITS copy of nsigma%f: hand over the host's current configuration
    call check(alf_b200_set_fields(alf_b200_handle_ptr, nsigma%f), __FILE__, __LINE__)
  end subroutine push_fields
  subroutine pull_fields()
    call check(alf_b200_get_fields(alf_b200_handle_ptr, nsigma%f), __FILE__, __LINE__)
  end subroutine pull_fields
  subroutine push_udv(which, udv)        ! which: 0 udvl, 1 udvr
    integer, intent(in) :: which
    class(UDV_State), intent(in) :: udv(:)
    integer :: nf_eff
    do nf_eff = 1, N_FL_eff
       if (allocated(udv(nf_eff)%V)) then
          call check(alf_b200_set_udv(alf_b200_handle_ptr, which, 0, 0, Calc_Fl_map(nf_eff), udv(nf_eff)%U, udv(nf_eff)%D, udv(nf_eff)%V), __FILE__, __LINE__)
       else
          call check(alf_b200_set_udv(alf_b200_handle_ptr, which, 0, 0, Calc_Fl_map(nf_eff), udv(nf_eff)%U, udv(nf_eff)%D, udv(nf_eff)%U), __FILE__, __LINE__)
       endif
    enddo
  end subroutine push_udv
  subroutine pull_udv(which, udv)
    integer, intent(in) :: which
    class(UDV_State), intent(inout) :: udv(:)
    complex(kind=dp), allocatable :: Utmp(:,:), Vtmp(:,:), Dtmp(:)
    integer :: nf_eff
    allocate(Utmp(Ndim,Ndim), Vtmp(Ndim,Ndim), Dtmp(Ndim))
    do nf_eff = 1, N_FL_eff
       call check(alf_b200_get_udv(alf_b200_handle_ptr, which, 0, 0, Calc_Fl_map(nf_eff), Utmp, Dtmp, Vtmp), __FILE__, __LINE__)
       udv(nf_eff)%U = Utmp(:, 1:udv(nf_eff)%n_part)
       udv(nf_eff)%D = Dtmp(1:udv(nf_eff)%n_part)
       if (allocated(udv(nf_eff)%V)) udv(nf_eff)%V = Vtmp(1:udv(nf_eff)%n_part, 1:udv(nf_eff)%n_part)
    enddo
    deallocate(Utmp, Vtmp, Dtmp)
  end subroutine pull_udv
  subroutine push_green(GR)
    complex(kind=dp), intent(in) :: GR(:,:,:)
    integer :: nf_eff
    do nf_eff = 1, N_FL_eff
       call check(alf_b200_set_green(alf_b200_handle_ptr, 0, Calc_Fl_map(nf_eff), GR(:,:,Calc_Fl_map(nf_eff))), __FILE__, __LINE__)
    enddo
  end subroutine push_green
  subroutine pull_green(GR)
    complex(kind=dp), intent(inout) :: GR(:,:,:)
    complex(kind=dp), allocatable :: G1(:,:)
    integer :: nf_eff
    allocate(G1(Ndim,Ndim))
    do nf_eff = 1, N_FL_eff
       call check(alf_b200_get_green(alf_b200_handle_ptr, 0, Calc_Fl_map(nf_eff), 0, G1), __FILE__, __LINE__)
       GR(:,:,Calc_Fl_map(nf_eff)) = G1
    enddo
    deallocate(G1)
  end subroutine pull_green

  ! ----------------------------------------------------------------------------------------------------------- reference procedures
  !> Prog/wrapur_mod.F90:37, 102-103.
  SUBROUTINE WRAPUR(NTAU, NTAU1, UDVR)
    CLASS(UDV_State), intent(inout), allocatable, dimension(:) :: UDVR
    Integer, Intent(IN) :: NTAU1, NTAU
    call push_fields(); call push_udv(1, UDVR)
    call check(alf_b200_wrapur(alf_b200_handle_ptr, NTAU, NTAU1), __FILE__, __LINE__)
    call pull_udv(1, UDVR)
  END SUBROUTINE WRAPUR

  !> Prog/wrapul_mod.F90:36, 108-110.
  SUBROUTINE WRAPUL(NTAU1, NTAU, UDVL)
    CLASS(UDV_State), intent(inout), allocatable, dimension(:) :: UDVL
    Integer, Intent(IN) :: NTAU1, NTAU
    call push_fields(); call push_udv(0, UDVL)
    call check(alf_b200_wrapul(alf_b200_handle_ptr, NTAU1, NTAU), __FILE__, __LINE__)
    call pull_udv(0, UDVL)
  END SUBROUTINE WRAPUL

  !> UDV_State%decompose, Prog/udv_state_mod.F90:448-452 (bound as `decompose => decompose_UDV_state`).
  SUBROUTINE decompose_UDV_state(UDVR)
    CLASS(UDV_State), intent(inout) :: UDVR
    ! stand-alone decompose of a projector state (U(Ndim, N_part), no V): WRAPUR / WRAPUL decompose such states on the device
    if (UDVR%n_part /= UDVR%ndim .or. .not. allocated(UDVR%V)) call Terminate_on_error(ERROR_GENERIC, __FILE__, __LINE__)
    call check(alf_b200_test_udv_decompose(alf_b200_device, 1, UDVR%ndim, 1, UDVR%side, UDVR%U, UDVR%D, UDVR%V), __FILE__, __LINE__)
  END SUBROUTINE decompose_UDV_state

  !> Prog/QDRP_decompose_mod.F90:60-69.  WORK / LWORK (LAPACK workspace, allocated by the reference and freed by its callers) are allocated with one
  !> element so that the callers' deallocate statements keep working.
  SUBROUTINE QDRP_decompose(Ndim_, N_part, Mat, D, IPVT, TAU, WORK, LWORK)
    Integer, intent(in) :: Ndim_
    Integer, intent(in) :: N_part
    Integer, intent(inout) :: LWORK
    Integer, Dimension(:), intent(inout), Allocatable :: IPVT
    COMPLEX(Kind=Kind(0.d0)), Dimension(:,:), Intent(inout) :: Mat
    COMPLEX(Kind=Kind(0.d0)), Dimension(:), Intent(inout) :: D
    COMPLEX(Kind=Kind(0.d0)), Dimension(:), Intent(inout), Allocatable :: TAU
    COMPLEX(Kind=Kind(0.d0)), Dimension(:), Intent(INOUT), Allocatable :: WORK
    complex(kind=dp), allocatable :: A1(:,:)
    real(kind=dp), allocatable :: D1(:), ph(:)
    integer(c_int), allocatable :: jp(:)
    allocate(A1(Ndim_, N_part), D1(N_part), jp(N_part), ph(5))
    if (.not. allocated(TAU)) allocate(TAU(N_part))
    A1 = Mat(1:Ndim_, 1:N_part)
    call check(alf_b200_test_qdrp(alf_b200_device, 1, Ndim_, N_part, 1, A1, D1, jp, TAU, ph), __FILE__, __LINE__)
    Mat(1:Ndim_, 1:N_part) = A1
    D(1:N_part) = cmplx(D1, 0.d0, kind=dp)
    IPVT(1:N_part) = jp
    if (.not. allocated(WORK)) allocate(WORK(1))
    LWORK = 1
    deallocate(A1, D1, jp, ph)
  END SUBROUTINE QDRP_decompose

  !> Prog/cgr1_mod.F90:36, 183-186.
  SUBROUTINE CGR(PHASE, NVAR, GRUP, udvr, udvl)
    CLASS(UDV_State), INTENT(IN) :: udvl, udvr
    COMPLEX(Kind=Kind(0.d0)), Dimension(:,:), Intent(INOUT) :: GRUP
    COMPLEX(Kind=Kind(0.d0)), Intent(INOUT) :: PHASE
    INTEGER         :: NVAR
    integer :: stab
    if (udvl%n_part < udvl%ndim) then          ! cgr1_mod.F90:207-211
       call CGRP(PHASE, GRUP, udvr, udvl)
       return
    endif
    stab = 0
#if defined(STAB3)
    stab = 3
#endif
    ! detUR, detUL absent (NULL): computed on the device
    call check(alf_b200_test_cgr(alf_b200_device, 1, udvl%ndim, 1, NVAR, stab, udvr%U, udvr%D, udvr%V, udvl%U, udvl%D, udvl%V, &
         & G=GRUP, phase=PHASE), __FILE__, __LINE__)
  END SUBROUTINE CGR

  !> Prog/cgr1_mod.F90:464-468.
  SUBROUTINE CGRP(phase, GRUP, udvr, udvl)
    CLASS(UDV_State), INTENT(IN) :: udvl, udvr
    COMPLEX (Kind=Kind(0.d0)), Dimension(:,:), Intent(OUT) :: GRUP
    COMPLEX (Kind=Kind(0.d0)), Intent(OUT) :: phase
    call check(alf_b200_test_cgrp(alf_b200_device, 1, udvl%ndim, udvl%n_part, 1, udvr%U, udvl%U, GRUP, phase), __FILE__, __LINE__)
  END SUBROUTINE CGRP

  !> Prog/cgr2_2_mod.F90:196, 322-324.  out4 of the C entry point = GRT0, GR00, GRTT, GR0T.
  SUBROUTINE CGR2_2(GRT0, GR00, GRTT, GR0T, udv2, udv1, LQ)
    Integer,  intent(in) :: LQ
    CLASS(UDV_State), intent(in) :: udv1, udv2
    Complex (Kind=Kind(0.d0)), intent(inout) :: GRT0(LQ,LQ), GR0T(LQ,LQ), GR00(LQ,LQ), GRTT(LQ,LQ)
    complex(kind=dp), allocatable :: out4(:,:,:)
    integer :: stab
    stab = 0
#if defined(STAB3)
    stab = 3
#endif
    allocate(out4(LQ,LQ,4))
    call check(alf_b200_test_cgr2_2(alf_b200_device, 1, LQ, 1, stab, udv2%U, udv2%D, udv2%V, udv1%U, udv1%D, udv1%V, out4), __FILE__, __LINE__)
    GRT0 = out4(:,:,1); GR00 = out4(:,:,2); GRTT = out4(:,:,3); GR0T = out4(:,:,4)
    deallocate(out4)
  END SUBROUTINE CGR2_2

  !> Prog/UDV_WRAP_mod.F90:125 (STAB1 / STAB2 builds call it from wrapur_mod.F90:96, wrapul_mod.F90:98-102, cgr1_mod.F90:115,137); NCON only triggers a
  !> diagnostic print in the reference.
  SUBROUTINE UDV_Wrap_Pivot(A, U, D, V, NCON, N1, N2)
    COMPLEX (Kind=Kind(0.d0)), intent(in),    dimension(:,:) :: A
    COMPLEX (Kind=Kind(0.d0)), intent(inout), dimension(:,:) :: U, V
    COMPLEX (Kind=Kind(0.d0)), intent(inout), dimension(:)   :: D
    Integer, intent(in) :: NCON, N1, N2
    complex(kind=dp) :: A1(N1,N2), U1(N1,N2), V1(N2,N2), D1(N2)
    A1 = A(1:N1,1:N2)
    call check(alf_b200_udv_wrap_pivot(alf_b200_device, 1, N1, N2, 1, A1, U1, D1, V1), __FILE__, __LINE__)
    U(1:N1,1:N2) = U1; V(1:N2,1:N2) = V1; D(1:N2) = D1
  END SUBROUTINE UDV_Wrap_Pivot

  !> Hop_mod multiplications, Prog/Hop_mod.F90:143-250: In <- e^{-dtau T} In etc.  `which` of alf_b200_hop_apply: 0 mmthr, 1 mmthr_m1, 2 mmthl,
  !> 3 mmthl_m1, 4 mmthlc.  (t: time slice of time-dependent hoppings, Prog/OpT_time_dependent.F90 -- not supported: the tables are static.)
  subroutine hop_apply_general(which, In, nf)
    integer, intent(in) :: which, nf
    complex(kind=dp), intent(inout) :: In(:,:)
    complex(kind=dp), allocatable :: A(:,:)
    if (size(In,1) /= Ndim .or. size(In,2) /= Ndim) call Terminate_on_error(ERROR_GENERIC, __FILE__, __LINE__)   ! rectangular operands: go through WRAPUR / WRAPUL
    allocate(A(Ndim,Ndim)); A = In
    call check(alf_b200_hop_apply(alf_b200_handle_ptr, which, nf, A), __FILE__, __LINE__)
    In = A; deallocate(A)
  end subroutine hop_apply_general
  Subroutine Hop_mod_mmthr(In, nf, t)
    Complex (Kind=Kind(0.d0)), intent(INOUT)  :: IN(:,:)
    Integer, intent(IN) :: nf, t
    call hop_apply_general(0, In, nf)
  end Subroutine Hop_mod_mmthr
  Subroutine Hop_mod_mmthr_m1(In, nf, t)
    Complex (Kind=Kind(0.d0)), intent(INOUT)  :: IN(:,:)
    Integer :: nf
    integer, intent(in) :: t
    call hop_apply_general(1, In, nf)
  end Subroutine Hop_mod_mmthr_m1
  Subroutine Hop_mod_mmthl(In, nf, t)
    Complex (Kind=Kind(0.d0)), intent(INOUT)  :: IN(:,:)
    Integer :: nf
    integer, intent(in) :: t
    call hop_apply_general(2, In, nf)
  end Subroutine Hop_mod_mmthl
  Subroutine Hop_mod_mmthl_m1(In, nf, t)
    Complex (Kind=Kind(0.d0)), intent(INOUT)  :: IN(:,:)
    Integer :: nf
    integer, intent(in) :: t
    call hop_apply_general(3, In, nf)
  end Subroutine Hop_mod_mmthl_m1
  Subroutine Hop_mod_mmthlc(In, nf, t)
    Complex (Kind=Kind(0.d0)), intent(INOUT)  :: IN(:,:)
    Integer :: nf
    integer, intent(in) :: t
    call hop_apply_general(4, In, nf)
  end Subroutine Hop_mod_mmthlc
  !> Prog/Hop_mod.F90:279-299: Out(:,:,nf) = e^{-dtau T/2} In(:,:,nf) e^{+dtau T/2}
  Subroutine Hop_mod_Symm(Out, In, t1, t2)
    COMPLEX (Kind=Kind(0.d0)), Dimension(:,:,:), Intent(Out):: Out
    COMPLEX (Kind=Kind(0.d0)), Dimension(:,:,:), Intent(IN):: In
    integer, intent(in) :: t1
    integer, optional, intent(in) :: t2
    complex(kind=dp), allocatable :: A(:,:)
    integer :: nf_eff, nf
    allocate(A(Ndim,Ndim))
    do nf_eff = 1, N_FL_eff
       nf = Calc_Fl_map(nf_eff)
       A = In(:,:,nf)
       call check(alf_b200_hop_apply(alf_b200_handle_ptr, 5, nf, A), __FILE__, __LINE__)
       Out(:,:,nf) = A
    enddo
    deallocate(A)
  end Subroutine Hop_mod_Symm

  !> Prog/Wrapgr_mod.F90:81-99.  Propose_S0, Nt_sequential_*, N_Global_tau are properties of the run that were handed over at attach time
  !> (alf_b200_set_s0_ising / alf_b200_set_global_tau_sampling); they are accepted here for signature compatibility.
  SUBROUTINE WRAPGRUP(GR, NTAU, PHASE, Propose_S0, Nt_sequential_start, Nt_sequential_end, N_Global_tau)
    COMPLEX (Kind=Kind(0.d0)), INTENT(INOUT), allocatable ::  GR(:,:,:)
    COMPLEX (Kind=Kind(0.d0)), INTENT(INOUT) ::  PHASE
    INTEGER, INTENT(IN) :: NTAU
    LOGICAL, INTENT(IN) :: Propose_S0
    INTEGER, INTENT(IN) :: Nt_sequential_start, Nt_sequential_end, N_Global_tau
    complex(kind=dp) :: ph(1)
    call push_fields(); call push_green(GR)
    call check(alf_b200_wrapgrup(alf_b200_handle_ptr, NTAU), __FILE__, __LINE__)
    call pull_green(GR); call pull_fields()
    call check(alf_b200_get_phase(alf_b200_handle_ptr, ph), __FILE__, __LINE__)
    PHASE = ph(1)
  END SUBROUTINE WRAPGRUP

  !> Prog/Wrapgr_mod.F90:160-180.
  SUBROUTINE WRAPGRDO(GR, NTAU, PHASE, Propose_S0, Nt_sequential_start, Nt_sequential_end, N_Global_tau)
    COMPLEX (Kind=Kind(0.d0)), INTENT(INOUT), allocatable ::  GR(:,:,:)
    COMPLEX (Kind=Kind(0.d0)), INTENT(INOUT) ::  PHASE
    INTEGER, INTENT(IN) :: NTAU
    LOGICAL, INTENT(IN) :: Propose_S0
    INTEGER, INTENT(IN) :: Nt_sequential_start, Nt_sequential_end, N_Global_tau
    complex(kind=dp) :: ph(1)
    call push_fields(); call push_green(GR)
    call check(alf_b200_wrapgrdo(alf_b200_handle_ptr, NTAU), __FILE__, __LINE__)
    call pull_green(GR); call pull_fields()
    call check(alf_b200_get_phase(alf_b200_handle_ptr, ph), __FILE__, __LINE__)
    PHASE = ph(1)
  END SUBROUTINE WRAPGRDO

  !> Prog/tau_m_mod.F90:56-64.  The device walks the time slices with ITS storage udvst (filled by the sweep that precedes TAU_M in main.F90:874-876) and
  !> accumulates the time-displaced lattice observables of Predefined_Obs_tau_* on the device (alf_b200_obs_tau_enable); Hamiltonians with their own
  !> ObserT read the matrices back per time point with alf_b200_taum_capture / alf_b200_get_taum.
  SUBROUTINE TAU_M(udvst, GR, PHASE, NSTM, NWRAP, STAB_NT, LOBS_ST, LOBS_EN)
    Integer, Intent(In) :: NSTM, NWRAP
    CLASS(UDV_State), Dimension(:,:), ALLOCATABLE, INTENT(IN) :: udvst
    Complex (Kind=Kind(0.d0)), Intent(in) :: GR(NDIM,NDIM,N_FL),  Phase
    Integer, Intent(In) :: STAB_NT(0:NSTM)
    Integer, Intent(In) :: LOBS_ST, LOBS_EN
    call push_fields(); call push_green(GR)
    call check(alf_b200_tau_m(alf_b200_handle_ptr), __FILE__, __LINE__)
  END SUBROUTINE TAU_M

end module alf_b200_shim
