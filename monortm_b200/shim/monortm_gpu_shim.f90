!  monortm_gpu_shim.f90 -- ISO_C_BINDING shim between the unchanged monoRTM Fortran host
!  (monortm.f90, monortm_sub.F90, lblatm.f90, lnfl_mod.f90) and libmonortm_b200.so.
!
!  It replaces the three calls of src/monortm.f90:557-574 (MODM, CALCTMR, RTM).  The host keeps
!  reading MONORTM.IN / MONORTM_PROF.IN / TAPE3 and writing MONORTM.OUT exactly as before.
!
!  Kinds: the parity build linuxGNUdbl compiles with -fdefault-real-8 -fdefault-integer-8
!  (build/makefile.common:195-198), so default REAL == c_double and default INTEGER == c_int64_t;
!  brd_mol_flg is integer*4 in lnfl_mod itself.  Build the host with the same flags and link
!  -lmonortm_b200 (see INTEGRATION.md).
!
!  NOTE: this file cannot be compiled in the authoring environment (no Fortran compiler); it is
!  delivered as source and its argument marshalling is exercised through the ctypes/C++ callers,
!  which pass the same column-major arrays.
module monortm_gpu_shim
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: mrtm_gpu_init, mrtm_gpu_stage_lines, mrtm_gpu_modm, mrtm_gpu_calctmr, mrtm_gpu_rtm, mrtm_gpu_done

  type(c_ptr), save :: ctx = c_null_ptr

  type, bind(c) :: mrtm_opts
     integer(c_int32_t) :: use_global_range = 0
     integer(c_int32_t) :: line_mode = 0
     real(c_double)     :: v1_global = 0, v2_global = 0
     integer(c_int64_t) :: iw0 = 0
     type(c_ptr)        :: sel_count = c_null_ptr
     type(c_ptr)        :: sel_hash = c_null_ptr
     type(c_ptr)        :: stream = c_null_ptr
     type(c_ptr)        :: xamnt = c_null_ptr        ! XAMNT of COMMON /PATHX/ (IXSECT=1), c_loc(XAMNT)
     integer(c_int64_t) :: ld_xamnt = 0              ! MX_XS
  end type mrtm_opts

  interface
     integer(c_int) function c_mrtm_init(device, ctx) bind(c, name="mrtm_init")
       import :: c_int, c_ptr
       integer(c_int), value :: device
       type(c_ptr) :: ctx
     end function
     integer(c_int) function c_mrtm_init_multi(device_mask, ctx) bind(c, name="mrtm_init_multi")
       import :: c_int, c_ptr, c_int64_t
       integer(c_int64_t), value :: device_mask          ! uint64_t bit mask, 0 = every visible GPU
       type(c_ptr) :: ctx
     end function
     integer(c_int) function c_mrtm_host_xsread(dir, ixmols, names, xv1, xv2, nreg, regs) bind(c, name="mrtm_host_xsread")
       import :: c_int, c_ptr, c_int64_t, c_double, c_char
       character(kind=c_char) :: dir(*), names(*)
       integer(c_int64_t), value :: ixmols
       real(c_double), value :: xv1, xv2
       integer(c_int64_t) :: nreg
       type(c_ptr) :: regs
     end function
     integer(c_int) function c_mrtm_stage_xsec(ctx, nreg, regs) bind(c, name="mrtm_stage_xsec")
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: ctx, regs
       integer(c_int64_t), value :: nreg
     end function
     subroutine c_mrtm_host_xs_free(regs, nreg) bind(c, name="mrtm_host_xs_free")
       import :: c_ptr, c_int64_t
       type(c_ptr), value :: regs
       integer(c_int64_t), value :: nreg
     end subroutine
     integer(c_int) function c_mrtm_free(ctx) bind(c, name="mrtm_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     function c_mrtm_last_error(ctx) bind(c, name="mrtm_last_error") result(p)
       import :: c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr) :: p
     end function
     integer(c_int) function c_mrtm_stage_lines(ctx, nblm, iim, iso, xnu0, deltnu, e, alps, alpf, x, xg, s0, &
          rmol, sdep, brd_flg, brd_tmp, brd_hw, brd_shft) bind(c, name="mrtm_stage_lines")
       import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t
       type(c_ptr), value :: ctx
       integer(c_int64_t) :: nblm(*), iso(*)
       integer(c_int64_t), value :: iim
       real(c_double) :: xnu0(*), deltnu(*), e(*), alps(*), alpf(*), x(*), xg(*), s0(*), rmol(*), sdep(*)
       integer(c_int32_t) :: brd_flg(*)
       real(c_double) :: brd_tmp(*), brd_hw(*), brd_shft(*)
     end function
     integer(c_int) function c_mrtm_modm(ctx, nwn, wn, dvset, nlay, p, t, clw, o, o_by_mol, oc, o_clw, odxsec, &
          nmol, wkl, wbrodl, sclcpl, sclhw, y0res, cntnm, ixsect, ibrd, scor, opts) bind(c, name="mrtm_modm")
       import :: c_int, c_ptr, c_int64_t, c_double
       type(c_ptr), value :: ctx
       integer(c_int64_t), value :: nwn, nlay, nmol, ixsect, ibrd
       real(c_double), value :: dvset, sclcpl, sclhw, y0res
       real(c_double) :: wn(*), p(*), t(*), clw(*), o(*), o_by_mol(*), oc(*), o_clw(*), odxsec(*)
       real(c_double) :: wkl(*), wbrodl(*), cntnm(7), scor(*)
       type(c_ptr), value :: opts
     end function
     integer(c_int) function c_mrtm_calctmr(ctx, nlayrs, nwn, wn, t, tz, o, tmr) bind(c, name="mrtm_calctmr")
       import :: c_int, c_ptr, c_int64_t, c_double
       type(c_ptr), value :: ctx
       integer(c_int64_t), value :: nlayrs, nwn
       real(c_double) :: wn(*), t(*), tz(*), o(*), tmr(*)
     end function
     integer(c_int) function c_mrtm_rtm(ctx, iout, irt, nwn, wn, nlay, t, tz, o, tmpsfc, rup, trtot, rdn, &
          reflc, emiss, rad, tb, idu) bind(c, name="mrtm_rtm")
       import :: c_int, c_ptr, c_int64_t, c_double
       type(c_ptr), value :: ctx
       integer(c_int64_t), value :: iout, irt, nwn, nlay, idu
       real(c_double) :: wn(*), t(*), tz(*), o(*), tmpsfc, rup(*), trtot(*), rdn(*), reflc(*), emiss(*), rad(*), tb(*)
     end function
  end interface

contains

  subroutine check(rc, where)
    integer(c_int), intent(in) :: rc
    character(*), intent(in) :: where
    if (rc /= 0) then
       write(*,*) 'monortm_b200: error ', rc, ' in ', where
       stop 'monortm_b200 GPU hot path failed'       ! the reference STOPs on every error path
    end if
  end subroutine check

  subroutine mrtm_gpu_init(device)
    integer, intent(in) :: device
    if (.not. c_associated(ctx)) call check(c_mrtm_init(int(device, c_int), ctx), 'mrtm_init')
  end subroutine

  !  every GPU of the node behind one context (bit d of mask = CUDA device d, 0 = all): the library splits each call
  subroutine mrtm_gpu_init_multi(mask)
    integer, intent(in) :: mask
    if (.not. c_associated(ctx)) call check(c_mrtm_init_multi(int(mask, c_int64_t), ctx), 'mrtm_init_multi')
  end subroutine

  !  IXSECT=1: call after XSREAD (src/monortm.f90:497).  The tables MONORTM_XSEC_SUB would re-read for every profile
  !  (src/monortm_sub.F90:1656-1671) are read once from the same FSCDXS / xs files in the working directory and staged.
  subroutine mrtm_gpu_stage_xsec(xv1, xv2)
    use lblparams, only: mx_xs, mxlay
    real*8, intent(in) :: xv1, xv2
    character*10 :: XSFILE, XSNAME, ALIAS
    common /XSECTF/ XSFILE(6,5,mx_xs), XSNAME(mx_xs), ALIAS(4,mx_xs)
    common /PATHX/ IXMAX, IXMOLS, IXINDX(mx_xs), XAMNT(mx_xs,mxlay)
    character(kind=c_char) :: names(10 * mx_xs)
    integer(c_int64_t) :: nreg
    type(c_ptr) :: regs
    integer :: i, j
    do i = 1, IXMOLS
       do j = 1, 10
          names(10 * (i - 1) + j) = XSNAME(i)(j:j)
       end do
    end do
    call check(c_mrtm_host_xsread('.' // c_null_char, int(IXMOLS, c_int64_t), names, xv1, xv2, nreg, regs), 'mrtm_host_xsread')
    call check(c_mrtm_stage_xsec(ctx, nreg, regs), 'mrtm_stage_xsec')
    call c_mrtm_host_xs_free(regs, nreg)
  end subroutine

  !  call once, right after GET_LNFL has filled the lnfl_mod module arrays (modm.f90:187-190)
  subroutine mrtm_gpu_stage_lines()
    use lnfl_mod, only: NBLM, ISO, XNU0, DELTNU, E, ALPS, ALPF, X, XG, S0, RMOL, SDEP, &
                        brd_mol_flg, brd_mol_tmp, brd_mol_hw, brd_mol_shft
    call check(c_mrtm_stage_lines(ctx, NBLM, int(size(XNU0, 2), c_int64_t), ISO, XNU0, DELTNU, E, ALPS, ALPF, X, XG, &
         S0, RMOL, SDEP, brd_mol_flg, brd_mol_tmp, brd_mol_hw, brd_mol_shft), 'mrtm_stage_lines')
  end subroutine

  !  drop-in for CALL MODM(...) (src/monortm.f90:557-561).  tips_2003 stays on this side.
  !  O, O_CLW, ODXSEC are (nwn, mxlay) and O_BY_MOL, OC (nwn, mxmol, mxlay) in the driver
  !  (src/monortm.f90:352-353): pass contiguous (nwn,nlay) / (nwn,mxmol,nlay) sections.
  subroutine mrtm_gpu_modm(nwn, wn, dvset, nlay, p, t, clw, o, o_by_mol, oc, o_clw, odxsec, nmol, wkl, wbrodl, &
       sclcpl, sclhw, y0res, cntnm7, ixsect, ibrd)
    integer, intent(in) :: nwn, nlay, nmol, ixsect, ibrd
    real*8, intent(in) :: wn(*)
    real, intent(in) :: dvset, p(*), t(*), clw(*), wkl(39, *), wbrodl(*), sclcpl, sclhw, y0res, cntnm7(7)
    real :: o(nwn, *), o_by_mol(nwn, 39, *), oc(nwn, 39, *), o_clw(nwn, *), odxsec(nwn, *)
    real, allocatable :: scor(:, :, :)
    integer :: k
    type(mrtm_opts), target :: opts
    real, target :: XAMNT
    common /PATHX/ IXMAX, IXMOLS, IXINDX(38), XAMNT(38, 603)      ! mx_xs, mxlay (src/lblparams.f90:29)
    if (ixsect == 1) then               ! MODM calls MONORTM_XSEC_SUB itself (modm.f90:197-198): XAMNT crosses through opts
       opts%xamnt = c_loc(XAMNT)
       opts%ld_xamnt = 38
    end if
    allocate(scor(42, 9, nlay))
    scor = 0.
    do k = 1, nlay
       call tips_2003(nmol, t(k), scor(:, :, k))          ! src/modm.f90:250
    end do
    call check(c_mrtm_modm(ctx, int(nwn, c_int64_t), wn, dvset, int(nlay, c_int64_t), p, t, clw, o, o_by_mol, oc, &
         o_clw, odxsec, int(nmol, c_int64_t), wkl, wbrodl, sclcpl, sclhw, y0res, cntnm7, &
         int(ixsect, c_int64_t), int(ibrd, c_int64_t), scor, c_loc(opts)), 'mrtm_modm')
    deallocate(scor)
  end subroutine

  subroutine mrtm_gpu_calctmr(nlayrs, nwn, wn, t, tz, o, tmr)
    integer, intent(in) :: nlayrs, nwn
    real*8, intent(in) :: wn(*)
    real, intent(in) :: t(*), tz(0:*), o(nwn, *)
    real :: tmr(*)
    call check(c_mrtm_calctmr(ctx, int(nlayrs, c_int64_t), int(nwn, c_int64_t), wn, t, tz, o, tmr), 'mrtm_calctmr')
  end subroutine

  subroutine mrtm_gpu_rtm(iout, irt, nwn, wn, nlay, t, tz, o, tmpsfc, rup, trtot, rdn, reflc, emiss, rad, tb, idu)
    integer, intent(in) :: iout, irt, nwn, nlay, idu
    real*8, intent(in) :: wn(*)
    real, intent(in) :: t(*), tz(0:*), o(nwn, *), reflc(*), emiss(*)
    real :: tmpsfc, rup(*), trtot(*), rdn(*), rad(*), tb(*)
    call check(c_mrtm_rtm(ctx, int(iout, c_int64_t), int(irt, c_int64_t), int(nwn, c_int64_t), wn, &
         int(nlay, c_int64_t), t, tz, o, tmpsfc, rup, trtot, rdn, reflc, emiss, rad, tb, int(idu, c_int64_t)), 'mrtm_rtm')
  end subroutine

  subroutine mrtm_gpu_done()
    if (c_associated(ctx)) call check(c_mrtm_free(ctx), 'mrtm_free')
    ctx = c_null_ptr
  end subroutine

end module monortm_gpu_shim
