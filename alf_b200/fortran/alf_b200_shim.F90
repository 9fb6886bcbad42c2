!  alf_b200_shim.F90 -- ISO_C_BINDING layer between ALF's unchanged Fortran (main.F90, Hamiltonian_main_mod,
!  Operator_mod, Fields_mod) and libalf_b200.so (include/alf_b200.h).
!
!  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (SURVEY.md F2).  It is the binding a
!  maintainer adds to Prog/ (see INTEGRATION.md); style follows ALF's only existing bind(c) code,
!  Libraries/Modules/lattices_interface_mod.F90:51-118 (value ints, assumed-size arrays, 1-based indices converted
!  on the C side).
module alf_b200_shim
  use iso_c_binding
  use runtime_error_mod          ! Terminate_on_error, ERROR_* codes (Libraries/Modules/runtime_error_mod.F90:56-130)
  use Operator_mod               ! type Operator (Prog/Operator_mod.F90:56-90)
  use Fields_mod                 ! type Fields   (Prog/Fields_mod.F90:79-99)
  implicit none
  private
  public :: alf_b200_attach, alf_b200_detach, alf_b200_batched_sweep, alf_b200_handle_ptr, UDV_Wrap_Pivot

  type(c_ptr), save :: alf_b200_handle_ptr = c_null_ptr
  integer(c_int), save :: alf_b200_device = 0      ! set by alf_b200_attach

  interface
     integer(c_int) function alf_b200_create(h, ndim, n_fl, n_sun, ltrot, nwrap, n_opv, n_opt, symm, stab, n_chains, device) &
          bind(c, name="alf_b200_create")
       import :: c_ptr, c_int
       type(c_ptr), intent(out) :: h
       integer(c_int), value :: ndim, n_fl, n_sun, ltrot, nwrap, n_opv, n_opt, symm, stab, n_chains, device
     end function
     integer(c_int) function alf_b200_destroy(h) bind(c, name="alf_b200_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function alf_b200_set_op_v(h, n, nf, nn, n_non_zero, diag, typ, P, U, E, g_re, g_im, a_re, a_im) &
          bind(c, name="alf_b200_set_op_v")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: n, nf, nn, n_non_zero, diag, typ
       integer(c_int), intent(in) :: P(*)
       complex(c_double_complex), intent(in) :: U(*)     ! interleaved (re,im) = what the C side reads as double[2*nn*nn]
       real(c_double), intent(in) :: E(*)
       real(c_double), value :: g_re, g_im, a_re, a_im
     end function
     integer(c_int) function alf_b200_set_op_t(h, nc, nf, nn, diag, P, U, E, g_re, g_im) bind(c, name="alf_b200_set_op_t")
       import :: c_ptr, c_int, c_double, c_double_complex
       type(c_ptr), value :: h
       integer(c_int), value :: nc, nf, nn, diag
       integer(c_int), intent(in) :: P(*)
       complex(c_double_complex), intent(in) :: U(*)
       real(c_double), intent(in) :: E(*)
       real(c_double), value :: g_re, g_im
     end function
     integer(c_int) function alf_b200_finalize_model(h) bind(c, name="alf_b200_finalize_model")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function alf_b200_set_seeds(h, seeds) bind(c, name="alf_b200_set_seeds")
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int32_t), intent(in) :: seeds(*)
     end function
     integer(c_int) function alf_b200_init_sweep(h) bind(c, name="alf_b200_init_sweep")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function alf_b200_sweep_host(h, n_sweeps, ltau, fields_in, fields_out, obs_out, control_out) &
          bind(c, name="alf_b200_sweep_host")
       import :: c_ptr, c_int, c_double_complex, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: n_sweeps, ltau
       complex(c_double_complex), intent(in)  :: fields_in(*)     ! [n + n_opv*(nt-1) + n_opv*ltrot*chain]  == nsigma%f(n,nt) per chain
       complex(c_double_complex), intent(out) :: fields_out(*)
       real(c_double), intent(out) :: obs_out(*), control_out(*)
     end function
     integer(c_int) function alf_b200_wrapur(h, ntau, ntau1) bind(c, name="alf_b200_wrapur")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: ntau, ntau1
     end function
     integer(c_int) function alf_b200_wrapul(h, ntau1, ntau) bind(c, name="alf_b200_wrapul")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: ntau1, ntau
     end function
     integer(c_int) function alf_b200_wrapgrup(h, ntau) bind(c, name="alf_b200_wrapgrup")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: ntau
     end function
     integer(c_int) function alf_b200_wrapgrdo(h, ntau) bind(c, name="alf_b200_wrapgrdo")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: ntau
     end function
     integer(c_int) function alf_b200_cgr(h, nvar) bind(c, name="alf_b200_cgr")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: nvar
     end function
     integer(c_int) function alf_b200_tau_m(h) bind(c, name="alf_b200_tau_m")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function alf_b200_set_projector(h, thtrot, n_part) bind(c, name="alf_b200_set_projector")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: thtrot, n_part
     end function
     integer(c_int) function alf_b200_set_trial_wf(h, nf, P_L, P_R) bind(c, name="alf_b200_set_trial_wf")
       import :: c_ptr, c_int, c_double_complex
       type(c_ptr), value :: h
       integer(c_int), value :: nf
       complex(c_double_complex), intent(in) :: P_L(*), P_R(*)      ! WF_L(nf)%P, WF_R(nf)%P  (Ndim x N_part)
     end function
     integer(c_int) function alf_b200_tau_p(h, nst_in) bind(c, name="alf_b200_tau_p")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: nst_in
     end function
     integer(c_int) function alf_b200_wrapgr_set_position(h, m) bind(c, name="alf_b200_wrapgr_set_position")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: m
     end function
     integer(c_int) function alf_b200_wrapgr_placegr(h, m1, ntau) bind(c, name="alf_b200_wrapgr_placegr")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: m1, ntau
     end function
     integer(c_int) function alf_b200_wrapgr_random_update(h, ntau, n_moves, maxlen, flip_length, flip_list, flip_value, &
          &                                                  t0_ratio, s0_ratio, accepted, place_to) bind(c, name="alf_b200_wrapgr_random_update")
       import :: c_ptr, c_int, c_double, c_double_complex, c_int8_t
       type(c_ptr), value :: h
       integer(c_int), value :: ntau, n_moves, maxlen, place_to
       integer(c_int), intent(in) :: flip_length(*), flip_list(*)     ! [chain][move], [chain][move][maxlen] (1-based operator indices)
       complex(c_double_complex), intent(in) :: flip_value(*)
       real(c_double), intent(in) :: t0_ratio(*), s0_ratio(*)
       integer(c_int8_t), intent(out) :: accepted(*)
     end function
     integer(c_int) function alf_b200_set_lattice(h, n_unit, norb, site_cell, site_orb, imj) bind(c, name="alf_b200_set_lattice")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: n_unit, norb
       integer(c_int), intent(in) :: site_cell(*), site_orb(*), imj(*)     ! List(:,1), List(:,2), Latt%imj (column-major, 1-based)
     end function
     integer(c_int) function alf_b200_obs_tau_enable(h, on) bind(c, name="alf_b200_obs_tau_enable")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
     end function
     integer(c_int) function alf_b200_obs_eq_enable(h, on) bind(c, name="alf_b200_obs_eq_enable")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
     end function
     integer(c_int) function alf_b200_get_obs_tau(h, acc, bg, cnt) bind(c, name="alf_b200_get_obs_tau")
       import :: c_ptr, c_int, c_double, c_double_complex
       type(c_ptr), value :: h
       complex(c_double_complex), intent(out) :: acc(*), bg(*)      ! Obs_Latt(imj, nt, no_I, no_J) per channel; Obs_Latt0
       real(c_double), intent(out) :: cnt(2)                        ! N, sum of signs
     end function
     integer(c_int) function alf_b200_get_obs_eq(h, acc, bg, cnt) bind(c, name="alf_b200_get_obs_eq")
       import :: c_ptr, c_int, c_double, c_double_complex
       type(c_ptr), value :: h
       complex(c_double_complex), intent(out) :: acc(*), bg(*)
       real(c_double), intent(out) :: cnt(2)
     end function
     integer(c_int) function alf_b200_get_green(h, chain, nf, symmetrize, gout) bind(c, name="alf_b200_get_green")
       import :: c_ptr, c_int, c_double_complex
       type(c_ptr), value :: h
       integer(c_int), value :: chain, nf, symmetrize
       complex(c_double_complex), intent(out) :: gout(*)
     end function
     integer(c_int) function alf_b200_set_green(h, chain, nf, gin) bind(c, name="alf_b200_set_green")
       import :: c_ptr, c_int, c_double_complex
       type(c_ptr), value :: h
       integer(c_int), value :: chain, nf
       complex(c_double_complex), intent(in) :: gin(*)
     end function
     integer(c_int) function alf_b200_get_phase(h, ph) bind(c, name="alf_b200_get_phase")
       import :: c_ptr, c_int, c_double_complex
       type(c_ptr), value :: h
       complex(c_double_complex), intent(out) :: ph(*)
     end function
     integer(c_int) function alf_b200_set_s0_ising(h, n_terms, op_start, term_start, entry_op, entry_dt, w, open_boundaries, propose_s0) &
          bind(c, name="alf_b200_set_s0_ising")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: n_terms, open_boundaries, propose_s0
       integer(c_int), intent(in) :: op_start(*), term_start(*), entry_op(*), entry_dt(*)
       real(c_double), intent(in) :: w(*)
     end function
     integer(c_int) function alf_b200_set_s0_gaussian(h, on) bind(c, name="alf_b200_set_s0_gaussian")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
     end function
     integer(c_int) function alf_b200_set_global_tau_sampling(h, nt_sequential_start, nt_sequential_end, n_global_tau) &
          bind(c, name="alf_b200_set_global_tau_sampling")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: nt_sequential_start, nt_sequential_end, n_global_tau
     end function
     integer(c_int) function alf_b200_set_global_move_tau_ising(h, n_sites, move_start, move_fields, n_terms, site_term_start, term_start, &
          entry_op, entry_dt, w, open_boundaries) bind(c, name="alf_b200_set_global_move_tau_ising")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: n_sites, n_terms, open_boundaries
       integer(c_int), intent(in) :: move_start(*), move_fields(*), site_term_start(*), term_start(*), entry_op(*), entry_dt(*)
       real(c_double), intent(in) :: w(*)
     end function
     integer(c_int) function alf_b200_langevin_update(h, delta_t, max_force, delta_t_running) bind(c, name="alf_b200_langevin_update")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), value :: delta_t, max_force
       real(c_double), intent(out) :: delta_t_running(*)
     end function
     integer(c_int) function alf_b200_langevin_forces(h, forces) bind(c, name="alf_b200_langevin_forces")
       import :: c_ptr, c_int, c_double_complex
       type(c_ptr), value :: h
       complex(c_double_complex), intent(out) :: forces(*)
     end function
     integer(c_int) function alf_b200_compute_fermion_det(h, log_abs_det, phase_det) bind(c, name="alf_b200_compute_fermion_det")
       import :: c_ptr, c_int, c_double, c_double_complex
       type(c_ptr), value :: h
       real(c_double), intent(out) :: log_abs_det(*)
       complex(c_double_complex), intent(out) :: phase_det(*)
     end function
     integer(c_int) function alf_b200_udv_wrap_pivot(device, is_complex, n1, n2, batch, A, U, D, V) bind(c, name="alf_b200_udv_wrap_pivot")
       import :: c_int, c_double_complex
       integer(c_int), value :: device, is_complex, n1, n2, batch
       complex(c_double_complex), intent(in) :: A(*)
       complex(c_double_complex), intent(out) :: U(*), D(*), V(*)
     end function
  end interface

contains

  subroutine check(rc, file, line)        ! the C-ABI never aborts: map its return code to ALF's convention
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: file
    integer, intent(in) :: line
    if (rc /= 0) call Terminate_on_error(int(rc), file, line)
  end subroutine

  !> Called once after ham%Ham_set (Prog/main.F90:318-319): flattens the PUBLIC fields of Op_V / Op_T
  !> (Prog/Operator_mod.F90:57-69) into the C tables.  M_exp/E_exp/ExpOpT_vec are private in ALF and are rebuilt by
  !> alf_b200_finalize_model exactly as Op_set/Op_exp/Hop_mod_init do.
  subroutine alf_b200_attach(Op_V, Op_T, Ndim, N_FL, N_SUN, Ltrot, Nwrap, Symm, n_chains, device, seeds)
    type(Operator), intent(in) :: Op_V(:,:), Op_T(:,:)
    integer, intent(in) :: Ndim, N_FL, N_SUN, Ltrot, Nwrap, n_chains, device
    logical, intent(in) :: Symm
    integer, intent(in) :: seeds(:)
    integer :: n, nf, stab
    alf_b200_device = device
    stab = 0
#if defined(STAB3)
    stab = 3
#endif
    call check(alf_b200_create(alf_b200_handle_ptr, Ndim, N_FL, N_SUN, Ltrot, Nwrap, size(Op_V,1), size(Op_T,1), &
         merge(1,0,Symm), stab, n_chains, device), __FILE__, __LINE__)
    do nf = 1, N_FL
       do n = 1, size(Op_V,1)
          call check(alf_b200_set_op_v(alf_b200_handle_ptr, n, nf, Op_V(n,nf)%N, Op_V(n,nf)%N_non_zero, merge(1,0,Op_V(n,nf)%diag), &
               Op_V(n,nf)%type, Op_V(n,nf)%P, Op_V(n,nf)%U, Op_V(n,nf)%E, dble(Op_V(n,nf)%g), aimag(Op_V(n,nf)%g), &
               dble(Op_V(n,nf)%alpha), aimag(Op_V(n,nf)%alpha)), __FILE__, __LINE__)
       enddo
       do n = 1, size(Op_T,1)
          call check(alf_b200_set_op_t(alf_b200_handle_ptr, n, nf, Op_T(n,nf)%N, merge(1,0,Op_T(n,nf)%diag), Op_T(n,nf)%P, &
               Op_T(n,nf)%U, Op_T(n,nf)%E, dble(Op_T(n,nf)%g), aimag(Op_T(n,nf)%g)), __FILE__, __LINE__)
       enddo
    enddo
    call check(alf_b200_finalize_model(alf_b200_handle_ptr), __FILE__, __LINE__)
    call check(alf_b200_set_seeds(alf_b200_handle_ptr, int(seeds, c_int32_t)), __FILE__, __LINE__)
  end subroutine

  !> Same name and argument list as Prog/UDV_WRAP_mod.F90:125 (STAB1 / STAB2 builds call it from wrapur_mod.F90:96,
  !> wrapul_mod.F90:98-102, cgr1_mod.F90:115,137); NCON only triggers a diagnostic print in the reference.
  subroutine UDV_Wrap_Pivot(A, U, D, V, NCON, N1, N2)
    complex(kind=kind(0.d0)), intent(in),    dimension(:,:) :: A
    complex(kind=kind(0.d0)), intent(inout), dimension(:,:) :: U, V
    complex(kind=kind(0.d0)), intent(inout), dimension(:)   :: D
    integer, intent(in) :: NCON, N1, N2
    complex(kind=kind(0.d0)) :: A1(N1,N2), U1(N1,N2), V1(N2,N2), D1(N2)
    A1 = A(1:N1,1:N2)
    call check(alf_b200_udv_wrap_pivot(alf_b200_device, 1, N1, N2, 1, A1, U1, D1, V1), __FILE__, __LINE__)
    U(1:N1,1:N2) = U1; V(1:N2,1:N2) = V1; D(1:N2) = D1
  end subroutine

  subroutine alf_b200_detach()
    integer(c_int) :: rc
    rc = alf_b200_destroy(alf_b200_handle_ptr); alf_b200_handle_ptr = c_null_ptr
  end subroutine

  !> Batched mode: replaces the body of the get_sequential() branch, Prog/main.F90:714-887 (+ TAU_M when Ltau == 1),
  !> for all chains of the handle.  nsigma_all(:,:,c) is chain c's nsigma%f.
  subroutine alf_b200_batched_sweep(nsigma_all, n_sweeps, ltau, obs, control)
    complex(kind=kind(0.d0)), intent(inout) :: nsigma_all(:,:,:)
    integer, intent(in) :: n_sweeps, ltau
    real(kind=kind(0.d0)), intent(out) :: obs(:), control(16)
    call check(alf_b200_sweep_host(alf_b200_handle_ptr, n_sweeps, ltau, nsigma_all, nsigma_all, obs, control), __FILE__, __LINE__)
  end subroutine

end module alf_b200_shim
