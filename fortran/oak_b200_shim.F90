! oak_b200_shim.F90 — ISO_C_BINDING layer between OAK's Fortran driver and liboak_b200.so.
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (it has no Fortran compiler); it is the binding a maintainer
! adds to OAK (see INTEGRATION.md).  It is deliberately thin and mechanical:
!
!   * interface blocks for the C entry points of include/oak_b200.h;
!   * oakb200_locanalysis(...): same argument list as locAnalysis (rrsqrt.F90:433-457) minus the
!     callback, which cannot cross a C ABI.  What the callback reads from module globals
!     (assimilation.F90:216-229,:3713-3767: zoneIndex, ModML, ModelGrid, obsGridX/Y/Z/T,
!     hCorrLengthToObs, hMaxCorrLengthToObs, loctype, metrictype) is flattened once per `init`
!     (zones) and once per `Assim` (observations);
!   * `class(Covar) R` is flattened with `select type` (DiagCovar -> D ; DCDCovar -> D and inner
!     DiagCovar; anything else -> ERROR_STOP: the GPU path supports diagonal R only).
!
! In assimilation.F90 the only change is, inside `Assim`, on the master thread between the barriers
! at :3215 and :3295 (the library is not re-entrant; under OpenMP all threads enter Assim):
!
!        if (schemetype.eq.LocalScheme) then
!   -        call locanalysis(zoneSize,selectObservations, xf,Hxf,yo,Sf,HSf, R, xa,Sa,locAmplitudes)
!   +  !$omp master
!   +        call oakb200_locanalysis(zoneSize, xf,Hxf,yo,Sf,HSf, R, xa,Sa,locAmplitudes)
!   +  !$omp end master
!   +  !$omp barrier
!
! Build OAK with PRECISION=double (real == real(c_double)); link with -loak_b200.

#include "ppdef.h"

module oak_b200
 use iso_c_binding
 implicit none
 private
 public :: oakb200_setup_zones, oakb200_locanalysis, oakb200_analysis, oakb200_shutdown

 type(c_ptr), save :: handle = c_null_ptr

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
     import; integer(c_int), value :: device; type(c_ptr) :: h; integer(c_int) :: rc
   end function
   function oakb200_destroy(h) bind(C, name='oakb200_destroy') result(rc)
     import; type(c_ptr), value :: h; integer(c_int) :: rc
   end function
   function oakb200_last_error() bind(C, name='oakb200_last_error') result(msg)
     import; type(c_ptr) :: msg
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
   function oakb200_set_observations(h, m, obsx, obsy, obsz, obst) bind(C, name='oakb200_set_observations') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: m
     real(c_double) :: obsx(*), obsy(*), obsz(*), obst(*)
     integer(c_int) :: rc
   end function
   ! table of the tabulated anamorphosis (AnamTrans%anam(v)%transform, K x 2), used by oakb200_assim_ensemble
   function oakb200_set_anamorphosis_table(h, K, table) bind(C, name='oakb200_set_anamorphosis_table') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int32_t), value :: K
     real(c_double) :: table(K,*)
     integer(c_int) :: rc
   end function
   function oakb200_local_analysis(h, n, N, m, xf, Hxf, yo, Sf, ldSf, HSf, ldHSf, Rdiag, d01, xa, Sa, ldSa, &
        amplitudes, stats) bind(C, name='oakb200_local_analysis') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: n, ldSf, ldHSf, ldSa
     integer(c_int32_t), value :: N, m
     real(c_double) :: xf(*), Hxf(*), yo(*), Sf(ldSf,*), HSf(ldHSf,*), Rdiag(*), xa(*), Sa(ldSa,*)
     type(c_ptr), value :: d01, amplitudes      ! optional arrays: c_null_ptr or c_loc(array)
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
   function oakb200_global_analysis(h, n, N, m, xf, Hxf, yo, Sf, ldSf, HSf, ldHSf, Rdiag, d01, xa, Sa, ldSa, &
        amplitudes, stats) bind(C, name='oakb200_global_analysis') result(rc)
     import
     type(c_ptr), value :: h
     integer(c_int64_t), value :: n, ldSf, ldHSf, ldSa
     integer(c_int32_t), value :: N, m
     real(c_double) :: xf(*), Hxf(*), yo(*), Sf(ldSf,*), HSf(ldHSf,*), Rdiag(*), xa(*), Sa(ldSa,*)
     type(c_ptr), value :: d01, amplitudes      ! optional arrays: c_null_ptr or c_loc(array)
     type(oakb200_stats) :: stats
     integer(c_int) :: rc
   end function
 end interface

contains

 subroutine check(rc, where)
  integer(c_int), intent(in) :: rc
  character(len=*), intent(in) :: where
  character(kind=c_char), pointer :: msg(:)
  if (rc /= 0) then
    call c_f_pointer(oakb200_last_error(), msg, [512])
    write(stderr,*) 'oak_b200: ', where, ' failed with status ', rc, ': ', msg(1:index(transfer(msg,repeat(' ',512)),c_null_char)-1)
    ERROR_STOP
  end if
 end subroutine

 ! Once per init(), after initPartition (assimilation.F90:400-430): position and localisation lengths
 ! of the FIRST element of every zone, exactly what selectObservations looks up for the index it is
 ! given (rrsqrt.F90:368 passes startIndex(zi); assimilation.F90:3713-3740,:3756,:3767).
 subroutine oakb200_setup_zones(device)
  use assimilation, only: zoneSize, zoneIndex, startIndexZones, ModML, ModelGrid, hCorrLengthToObs, &
       hMaxCorrLengthToObs, loctype, metrictype, ind2submv
  integer, intent(in) :: device
  integer :: zi, nz, index, v, i, j, k, n
  logical :: out
  real(c_double), allocatable :: zx(:), zy(:), zz(:), zt(:), cl(:), ml(:)
  real :: x4(4)

  nz = size(zoneSize)
  allocate(zx(nz), zy(nz), zz(nz), zt(nz), cl(nz), ml(nz))
  zx = 0; zy = 0; zz = 0; zt = 0
  do zi = 1, nz
    index = zoneIndex(startIndexZones(zi))
    call ind2submv(ModML, index, v, i, j, k, n)
    select case (ModML%ndim(v))
    case (1); x4(1:1) = ModelGrid(v)%getCoord((/ i /), out)
    case (2); x4(1:2) = ModelGrid(v)%getCoord((/ i,j /), out)
    case (3); x4(1:3) = ModelGrid(v)%getCoord((/ i,j,k /), out)
    case (4); x4(1:4) = ModelGrid(v)%getCoord((/ i,j,k,n /), out)
    end select
    zx(zi) = x4(1)
    if (ModML%ndim(v) >= 2) zy(zi) = x4(2)
    if (ModML%ndim(v) >= 3) zz(zi) = x4(3)
    if (ModML%ndim(v) >= 4) zt(zi) = x4(4)
    cl(zi) = hCorrLengthToObs(index)
    ml(zi) = hMaxCorrLengthToObs(index)
  end do
  if (.not. c_associated(handle)) call check(oakb200_create(int(device,c_int), handle), 'oakb200_create')
  ! weightfun 0 = the Gaussian callback of assimilation.F90:3767
  call check(oakb200_set_zones(handle, int(nz,c_int32_t), int(zoneSize,c_int32_t), zx, zy, zz, zt, cl, ml, &
       int(loctype,c_int32_t), int(metrictype,c_int32_t), 0_c_int32_t), 'oakb200_set_zones')
 end subroutine

 ! Drop-in for: call locanalysis(zoneSize,selectObservations,xf,Hxf,yo,Sf,HSf,R,xa,Sa,locAmplitudes)
 subroutine oakb200_locanalysis(zoneSize, xf, Hxf, yo, Sf, HSf, R, xa, Sa, amplitudes)
  use covariance
  use assimilation, only: obsGridX, obsGridY, obsGridZ, obsGridT
  integer, intent(in) :: zoneSize(:)
  real(c_double), intent(in) :: xf(:), Hxf(:), yo(:), Sf(:,:), HSf(:,:)
  class(Covar), intent(in) :: R
  real(c_double), intent(out) :: xa(:)
  real(c_double), intent(out), target :: Sa(:,:)
  real(c_double), intent(out), optional, target :: amplitudes(:,:)
  real(c_double), allocatable, target :: Rdiag(:), d01(:)
  type(c_ptr) :: pd01, pamp
  type(oakb200_stats) :: stats
  integer :: m

  m = size(yo)
  pd01 = c_null_ptr
  select type (R)
  type is (DiagCovar)                    ! covariance.F90:70-79 ; built at assimilation.F90:2120-2127
    allocate(Rdiag(m)); Rdiag = R%D
  type is (DCDCovar)                     ! covariance.F90:109-118 ; excluded observations, assimilation.F90:3086-3092
    allocate(Rdiag(m), d01(m)); d01 = R%D
    select type (C => R%C)
    type is (DiagCovar)
      Rdiag = C%D
    class default
      write(stderr,*) 'oak_b200: DCDCovar with a non-diagonal inner covariance is not supported'; ERROR_STOP
    end select
    pd01 = c_loc(d01)
  class default
    write(stderr,*) 'oak_b200: only diagonal observation error covariances are supported'; ERROR_STOP
  end select
  pamp = c_null_ptr
  if (present(amplitudes)) pamp = c_loc(amplitudes)

  ! observation positions change with every Assim call (assimilation.F90:3155-3160)
  call check(oakb200_set_observations(handle, int(m,c_int32_t), obsGridX, obsGridY, obsGridZ, obsGridT), &
       'oakb200_set_observations')
  call check(oakb200_local_analysis(handle, int(size(xf),c_int64_t), int(size(Sf,2),c_int32_t), int(m,c_int32_t), &
       xf, Hxf, yo, Sf, int(size(Sf,1),c_int64_t), HSf, int(size(HSf,1),c_int64_t), Rdiag, pd01, &
       xa, Sa, int(size(Sa,1),c_int64_t), pamp, stats), 'oakb200_local_analysis')
 end subroutine

 ! Drop-in for the global scheme: call analysis(xf,Hxf,yo,Sf,HSf,R,xa,Sa,amplitudes)   (rrsqrt.F90:196-208;
 ! the schemetype = 0 branch of Assim)
 subroutine oakb200_analysis(xf, Hxf, yo, Sf, HSf, R, xa, Sa, amplitudes)
  use covariance
  real(c_double), intent(in) :: xf(:), Hxf(:), yo(:), Sf(:,:), HSf(:,:)
  class(Covar), intent(in) :: R
  real(c_double), intent(out) :: xa(:)
  real(c_double), intent(out), target :: Sa(:,:)
  real(c_double), intent(out), optional, target :: amplitudes(:)
  real(c_double), allocatable, target :: Rdiag(:), d01(:)
  type(c_ptr) :: pd01, pamp
  type(oakb200_stats) :: stats
  integer :: m

  m = size(yo)
  pd01 = c_null_ptr
  select type (R)
  type is (DiagCovar)
    allocate(Rdiag(m)); Rdiag = R%D
  type is (DCDCovar)
    allocate(Rdiag(m), d01(m)); d01 = R%D
    select type (C => R%C)
    type is (DiagCovar)
      Rdiag = C%D
    class default
      write(stderr,*) 'oak_b200: DCDCovar with a non-diagonal inner covariance is not supported'; ERROR_STOP
    end select
    pd01 = c_loc(d01)
  class default
    write(stderr,*) 'oak_b200: only diagonal observation error covariances are supported'; ERROR_STOP
  end select
  pamp = c_null_ptr
  if (present(amplitudes)) pamp = c_loc(amplitudes)
  call check(oakb200_global_analysis(handle, int(size(xf),c_int64_t), int(size(Sf,2),c_int32_t), int(m,c_int32_t), &
       xf, Hxf, yo, Sf, int(size(Sf,1),c_int64_t), HSf, int(size(HSf,1),c_int64_t), Rdiag, pd01, &
       xa, Sa, int(size(Sa,1),c_int64_t), pamp, stats), 'oakb200_global_analysis')
 end subroutine

 subroutine oakb200_shutdown()
  integer(c_int) :: rc
  if (c_associated(handle)) rc = oakb200_destroy(handle)
  handle = c_null_ptr
 end subroutine

end module oak_b200
