! oak_b200_iface.F90 — ISO_C_BINDING interfaces of liboak_b200.so (include/oak_b200.h).
!
! This module depends on NOTHING of OAK (only iso_c_binding), so `module assimilation` can `use oak_b200_iface`
! without a circular dependency; the wrappers that need OAK's module globals (zoneIndex, ModML, ModelGrid,
! obsGridX..T, hCorrLengthToObs, ...) are NOT here but in oak_b200_assim.inc, which is #include'd in the
! `contains` section of module assimilation itself (INTEGRATION.md section 2).
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is installed).  Fortran is case-insensitive:
! the dummies of the analysis entry points are therefore nrows / nens / nobs, not n / N / m.
! Build OAK with PRECISION=double (real == real(c_double)); link with -loak_b200.
module oak_b200_iface
 use iso_c_binding
 implicit none
 public

 integer, parameter :: OAKB200_MAX_PEERS = 16

 type, bind(C) :: oakb200_stats
   integer(c_int64_t) :: zones_total, zones_skipped, obs_relevant_sum, obs_candidate_sum, jacobi_sweeps_sum
   integer(c_int64_t) :: h2d_bytes, d2h_bytes
   real(c_double)     :: ms_total, ms_pack, ms_gram, ms_eig, ms_apply
   integer(c_int64_t) :: launches
   integer(c_int64_t) :: zones_fallback
   real(c_double)     :: ms_tridiag, ms_tql, ms_tvec
 end type

 interface
   function oakb200_create(device, h) bind(C, name='oakb200_create') result(rc)
     import
     integer(c_int), value :: device
     type(c_ptr) :: h
     integer(c_int) :: rc
   end function
   function oakb200_destroy(h) bind(C, name='oakb200_destroy') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int) :: rc
   end function
   function oakb200_last_error() bind(C, name='oakb200_last_error') result(msg)
     import
     type(c_ptr) :: msg
   end function
   function oakb200_set_option(h, key, val) bind(C, name='oakb200_set_option') result(rc)
     import
     type(c_ptr), value :: h
     character(kind=c_char) :: key(*)          ! NUL-terminated: 'localise_obs'//c_null_char
     real(c_double), value :: val
     integer(c_int) :: rc
   end function
   ! page-locked host memory for Sf / Sa / HSf: call c_f_pointer(ptr, Sf, [n, N]) on the result
   function oakb200_host_alloc(bytes, ptr) bind(C, name='oakb200_host_alloc') result(rc)
     import
     integer(c_int64_t), value :: bytes
     type(c_ptr) :: ptr
     integer(c_int) :: rc
   end function
   function oakb200_host_free(ptr) bind(C, name='oakb200_host_free') result(rc)
     import
     type(c_ptr), value :: ptr
     integer(c_int) :: rc
   end function
   function oakb200_set_zones(h, nzones, zoneSize, zx, zy, zz, zt, corrLen, maxLen, loctype, metrictype, &
        weightfun) bind(C, name='oakb200_set_zones') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: nzones, loctype, metrictype, weightfun
     integer(c_int32_t) :: zoneSize(*)
     real(c_double) :: zx(*), zy(*), zz(*), zt(*), corrLen(*), maxLen(*)
     integer(c_int) :: rc
   end function
   function oakb200_set_observations(h, nobs, obsx, obsy, obsz, obst) &
        bind(C, name='oakb200_set_observations') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: nobs
     real(c_double) :: obsx(*), obsy(*), obsz(*), obst(*)
     integer(c_int) :: rc
   end function
   ! table of a tabulated anamorphosis valid for the whole state vector (AnamTrans%anam(v)%transform, K x 2)
   function oakb200_set_anamorphosis_table(h, K, table) bind(C, name='oakb200_set_anamorphosis_table') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: K
     real(c_double) :: table(*)
     integer(c_int) :: rc
   end function
   ! per-variable anamorphosis: what anamtransform looks up through ind2submv (assimilation.F90:4531-4567)
   function oakb200_set_anamorphosis_vars(h, nvar, vtype, vK, tables, nrows, rowvar) &
        bind(C, name='oakb200_set_anamorphosis_vars') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: nvar
     integer(c_int32_t) :: vtype(*), vK(*), rowvar(*)
     real(c_double) :: tables(*)
     integer(c_int64_t), value :: nrows
     integer(c_int) :: rc
   end function
   ! locanalysis / analysis with host arrays (rrsqrt.F90:433-466, :196-208)
   function oakb200_local_analysis(h, nrows, nens, nobs, xf, Hxf, yo, Sf, ldSf, HSf, ldHSf, Rdiag, d01, xa, Sa, &
        ldSa, amplitudes, stats) bind(C, name='oakb200_local_analysis') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: nrows, ldSf, ldHSf, ldSa
     integer(c_int32_t), value :: nens, nobs
     real(c_double) :: xf(*), Hxf(*), yo(*), Sf(ldSf,*), HSf(ldHSf,*), Rdiag(*), xa(*), Sa(ldSa,*)
     type(c_ptr), value :: d01, amplitudes      ! optional arrays: c_null_ptr or c_loc(array)
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
   function oakb200_global_analysis(h, nrows, nens, nobs, xf, Hxf, yo, Sf, ldSf, HSf, ldHSf, Rdiag, d01, xa, Sa, &
        ldSa, amplitudes, stats) bind(C, name='oakb200_global_analysis') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: nrows, ldSf, ldHSf, ldSa
     integer(c_int32_t), value :: nens, nobs
     real(c_double) :: xf(*), Hxf(*), yo(*), Sf(ldSf,*), HSf(ldHSf,*), Rdiag(*), xa(*), Sa(ldSa,*)
     type(c_ptr), value :: d01, amplitudes
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
   ! ensemble branch of Assim in one call (assimilation.F90:3106-3134, :3235 or :3288, :3301-3357, :3558-3562)
   function oakb200_assim_ensemble(h, nrows, nens, nobs, E, ldE, nnz, Hi, Hj, Hs, Hshift, yo, Rdiag, d01, &
        anamtype, inflation, maxCorrection, Ea, ldEa, xf_out, xa_out, stats) &
        bind(C, name='oakb200_assim_ensemble') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: nrows, ldE, nnz, ldEa
     integer(c_int32_t), value :: nens, nobs, anamtype
     real(c_double) :: E(ldE,*), Hs(*), yo(*), Rdiag(*), Ea(ldEa,*)
     integer(c_int32_t) :: Hi(*), Hj(*)          ! 1-based COO indices (matoper.F90:30-39)
     type(c_ptr), value :: Hshift, d01, maxCorrection, xf_out, xa_out   ! optional arrays
     real(c_double), value :: inflation
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
   ! relevant observations per zone as counted by the production kernel of the last analysis (diagnostics)
   function oakb200_zone_counts(h, mloc) bind(C, name='oakb200_zone_counts') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t) :: mloc(*)              ! one entry per zone
     integer(c_int) :: rc
   end function
   ! observation operator: batched cinterp (ndgrid.F90:1183-1257) for one model grid with separable axes; called from
   ! genObservationOper (assimilation.F90:2569-2585) once per model variable instead of once per observation
   function oakb200_cinterp(h, ndim, gshape, axes, masked, nobs, xi, indexes, coeff, nbp, ndegenerate) &
        bind(C, name='oakb200_cinterp') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: ndim, nobs
     integer(c_int32_t) :: gshape(*)            ! ndim entries
     real(c_double) :: axes(*)                  ! the coordinate axes one after the other
     type(c_ptr), value :: masked               ! c_loc of an integer(c_int8_t) array, 1 = masked, or c_null_ptr
     real(c_double) :: xi(ndim, *)              ! xi(:,l) = position of observation l
     integer(c_int32_t) :: indexes(ndim, 2**ndim, *)   ! 1-based corner subscripts: tmpHindex(7:6+ndim, ...)
     real(c_double) :: coeff(2**ndim, *)        ! tmpHcoeff
     integer(c_int32_t) :: nbp(*), ndegenerate  ! nbp = 2**ndim, 0 (out of grid / masked corner), -1 (degenerate cell)
     integer(c_int) :: rc
   end function
   ! multi-GPU (replaces parallPartion / parallGather, parall.F90:166-186, :507-566)
   function oakb200_partition_zones(nzones, nranks, first) bind(C, name='oakb200_partition_zones') result(rc)
     import
     integer(c_int32_t), value :: nzones, nranks
     integer(c_int32_t) :: first(*)             ! nranks+1 entries, 0-based first zone of every rank
     integer(c_int) :: rc
   end function
   function oakb200_ipc_alloc(h, bytes, ptr, ipchandle) bind(C, name='oakb200_ipc_alloc') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: bytes
     type(c_ptr) :: ptr
     character(kind=c_char) :: ipchandle(64)
     integer(c_int) :: rc
   end function
   function oakb200_ipc_open(h, ipchandle, ptr) bind(C, name='oakb200_ipc_open') result(rc)
     import
     type(c_ptr), value :: h
     character(kind=c_char) :: ipchandle(64)
     type(c_ptr) :: ptr
     integer(c_int) :: rc
   end function
   function oakb200_ipc_close(h, ptr) bind(C, name='oakb200_ipc_close') result(rc)
     import
     type(c_ptr), value :: h, ptr
     integer(c_int) :: rc
   end function
   function oakb200_ipc_free(h, ptr) bind(C, name='oakb200_ipc_free') result(rc)
     import
     type(c_ptr), value :: h, ptr
     integer(c_int) :: rc
   end function
   function oakb200_set_peer_outputs(h, npeer, Sa_peer, xa_peer, ld_peer, row0) &
        bind(C, name='oakb200_set_peer_outputs') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: npeer
     type(c_ptr) :: Sa_peer(*), xa_peer(*)       ! device pointers of every rank's result arrays
     integer(c_int64_t), value :: ld_peer, row0
     integer(c_int) :: rc
   end function
   ! device-pointer flavour: used by the multi-GPU wrapper (state slab resident on the device)
   function oakb200_local_analysis_dev(h, nrows, nens, nobs, xf, Hxf, yo, Sf, ldSf, HSf, ldHSf, Rdiag, d01, xa, Sa, &
        ldSa, amplitudes, stream, stats) bind(C, name='oakb200_local_analysis_dev') result(rc)
     import
     type(c_ptr), value :: h, xf, Hxf, yo, Sf, HSf, Rdiag, d01, xa, Sa, amplitudes, stream
     integer(c_int64_t), value :: nrows, ldSf, ldHSf, ldSa
     integer(c_int32_t), value :: nens, nobs
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
 end interface

contains

 ! non-zero status -> the reference's error convention (message on unit 0, exit(1); ppdef.h:22)
 subroutine oakb200_check(rc, where)
  integer(c_int), intent(in) :: rc
  character(len=*), intent(in) :: where
  character(kind=c_char), pointer :: msg(:)
  character(len=512) :: text
  integer :: i
  if (rc /= 0) then
    call c_f_pointer(oakb200_last_error(), msg, [512])
    text = ' '
    do i = 1, 512
      if (msg(i) == c_null_char) exit
      text(i:i) = msg(i)
    end do
    write(0,*) 'oak_b200: ', where, ' failed with status ', rc, ': ', trim(text)
    call exit(1)
  end if
 end subroutine

end module oak_b200_iface
