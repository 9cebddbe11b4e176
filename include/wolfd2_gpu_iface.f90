!-----------------------------------------------------------------------
! wolfd2_gpu_iface.f90 -- ISO_C_BINDING interface block for the residency
! API of libwolfd2_b200.so (include/wolfd2_b200.h, section (2)).
!
! UNTESTED HERE: neither the build container nor the GPU box has a Fortran
! front-end (gfortran/flang/nvfortran are all absent), so this file has not
! been compiled.  It is the binding a wolfd2 maintainer adds to the host
! program; see INTEGRATION.md for the two edits to src/main.f.
!
! The literal shims (nauxmomentum_, ppe_, project_, velboundcond_, ...) need
! NO interface block: they carry the gfortran F77 calling convention of the
! subroutines they replace (momentum.f:33, pressure.f:30, utility.f:253,
! bound_cond.f:511/853/1656), so main.f links against them unchanged once the
! original objects' symbols are removed from the link line.
!-----------------------------------------------------------------------
module wolfd2_gpu
  use, intrinsic :: iso_c_binding
  implicit none

  ! struct wolfd2_params  (include/wolfd2_b200.h)
  type, bind(C) :: w2_params
    integer(c_int32_t) :: nx, ny
    integer(c_int32_t) :: mqiter, nmeiter, nPpeSolver, msorit
    integer(c_int32_t) :: lCartesGrid, nfiltu, nfiltv, reserved
    real(c_double)     :: dk, re, fr, qtol, sortol, sorrel, fpu, fpv
  end type w2_params

  ! struct wolfd2_regions: addresses of the tables SetUpBCs built
  type, bind(C) :: w2_regions
    type(c_ptr) :: nReg, nRegBrd, nRegType, nMomBdTp
    type(c_ptr) :: dBCVal, dPRporos, dPRporc1, dPRporc2
  end type w2_regions

  ! struct wolfd2_metrics: addresses of the 30 metric arrays Grid built
  type, bind(C) :: w2_metrics
    type(c_ptr) :: rau, rbu, rbv, rgv, ran, rbn, rgn, rac, rbc, rgc
    type(c_ptr) :: dju, djv, djc, djn
    type(c_ptr) :: xen, yen, xzn, yzn, xec, yec, xzc, yzc
    type(c_ptr) :: xeu, yeu, xzv, yzv, xzu, yzu, xev, yev
  end type w2_metrics

  ! struct wolfd2_step_log: the PrintDiff tuple (string.f:547-559)
  type, bind(C) :: w2_step_log
    integer(c_int32_t) :: nQLiter, nSorConv, sor_converged, diverged
    real(c_double)     :: dif(4)
  end type w2_step_log

  ! struct wolfd2_thermal: thermal energy equation (thermal.f, main.f:840-894)
  type, bind(C) :: w2_thermal
    integer(c_int32_t) :: nthermen, neqstate, nfiltt, reserved
    real(c_double)     :: pe, dmeittol, fpt, uref, densref, tmax, tref, rconst
    type(c_ptr)        :: nTRgType, nTemBdTp, dTRgVal, dHGSTval
  end type w2_thermal

  ! struct wolfd2_smallscale: ATD small-scale model (small_scale.f:31-49, main.f:647-665)
  type, bind(C) :: w2_smallscale
    integer(c_int32_t) :: nsmallscl, nssPpeSlvr, mssSorIt, reserved
    real(c_double)     :: dlref, uref, tref, tmax
    real(c_double)     :: pe
    real(c_double)     :: ssSorTol, ssSorRel
    real(c_double)     :: ssFiltPar(4)
    real(c_double)     :: ssCu0, ssTsCoef, ssHsCoef, ssTemCoef, ssBnCrit, ssRMpMax, ssRMpExp
  end type w2_smallscale

  ! struct wolfd2_traject: scalar arguments of Traject (traject.f:154-164)
  type, bind(C) :: w2_traject
    integer(c_int32_t) :: ntr, ntsubstp, nTrMethod, nTrCdEq, mTrHTmit, reserved
    real(c_double)     :: densref, dTrHTtol, dTrHTdel
  end type w2_traject

  integer(c_int32_t), parameter :: W2_F_U = 0, W2_F_V = 1, W2_F_P = 2, W2_F_D = 8, W2_F_T = 11
  integer(c_int32_t), parameter :: W2_F_USS = 14, W2_F_VSS = 15, W2_F_PSS = 16, W2_F_TSS = 17

  interface
    ! replaces the compile-time include/config.f:19-26
    integer(c_int) function wolfd2_b200_config(mnx, mny, mgri, mgrj) bind(C, name='wolfd2_b200_config')
      import :: c_int, c_int32_t
      integer(c_int32_t), value :: mnx, mny, mgri, mgrj
    end function

    ! after Grid / SetUpBCs (main.f:434-448): upload metrics and region tables
    integer(c_int) function wolfd2_b200_create(ctx, par, reg, met) bind(C, name='wolfd2_b200_create')
      import :: c_int, c_ptr, w2_params, w2_regions, w2_metrics
      type(c_ptr), intent(out) :: ctx
      type(w2_params),  intent(in) :: par
      type(w2_regions), intent(in) :: reg
      type(w2_metrics), intent(in) :: met
    end function

    subroutine wolfd2_b200_destroy(ctx) bind(C, name='wolfd2_b200_destroy')
      import :: c_ptr
      type(c_ptr), value :: ctx
    end subroutine

    ! host (0:mnx,0:mny) array -> device field, and back (dumps, restart, ATD, trajectories)
    integer(c_int) function wolfd2_b200_upload_field(ctx, which, host) bind(C, name='wolfd2_b200_upload_field')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: which
      real(c_double), intent(in) :: host(*)
    end function
    integer(c_int) function wolfd2_b200_download_field(ctx, which, host) bind(C, name='wolfd2_b200_download_field')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: which
      real(c_double), intent(out) :: host(*)
    end function

    ! main.f:606-641
    integer(c_int) function wolfd2_b200_coldstart(ctx, nSorConv) bind(C, name='wolfd2_b200_coldstart')
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: nSorConv
    end function

    ! main.f:690-981, nsteps times, fields resident on the device
    integer(c_int) function wolfd2_b200_step(ctx, nsteps, logs) bind(C, name='wolfd2_b200_step')
      import :: c_int, c_int32_t, c_ptr, w2_step_log
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: nsteps
      type(w2_step_log), intent(out) :: logs(*)
    end function

    ! same with u,v,p crossing the boundary in host arrays every call
    integer(c_int) function wolfd2_b200_step_host(ctx, nsteps, u, v, p, logs) bind(C, name='wolfd2_b200_step_host')
      import :: c_int, c_int32_t, c_ptr, c_double, w2_step_log
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: nsteps
      real(c_double), intent(inout) :: u(*), v(*), p(*)
      type(w2_step_log), intent(out) :: logs(*)
    end function

    ! thermal energy equation on / off for the following steps
    integer(c_int) function wolfd2_b200_set_thermal(ctx, th) bind(C, name='wolfd2_b200_set_thermal')
      import :: c_int, c_ptr, w2_thermal
      type(c_ptr), value :: ctx
      type(w2_thermal), intent(in) :: th
    end function

    ! ATD small-scale model on / off (main.f:706-727, :896-940 then run inside wolfd2_b200_step)
    integer(c_int) function wolfd2_b200_set_smallscale(ctx, ss) bind(C, name='wolfd2_b200_set_smallscale')
      import :: c_int, c_ptr, w2_smallscale
      type(c_ptr), value :: ctx
      type(w2_smallscale), intent(in) :: ss
    end function
    ! main.f:643-665 (SmallScale with initflg = 0 on the uploaded fields)
    integer(c_int) function wolfd2_b200_smallscale_init(ctx) bind(C, name='wolfd2_b200_smallscale_init')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    ! one plane (1..3) of the saved chaotic-map iterates, family 0/1/2 = umap/vmap/tmap (restart files)
    integer(c_int) function wolfd2_b200_smallscale_map(ctx, family, plane, host, upload) &
        bind(C, name='wolfd2_b200_smallscale_map')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: family, plane, upload
      real(c_double), intent(inout) :: host(*)
    end function

    ! Lagrangian particles: main.f:1000-1024 (VelAvg, PTDAvg, Traject) then ends every step on the device
    integer(c_int) function wolfd2_b200_set_trajectories(ctx, tr, x, y, cpartx, cparty, repc, xp, yp, up, vp, nTOutBnd) &
        bind(C, name='wolfd2_b200_set_trajectories')
      import :: c_int, c_int32_t, c_ptr, c_double, w2_traject
      type(c_ptr), value :: ctx
      type(w2_traject), intent(in) :: tr
      real(c_double), intent(in) :: x(*), y(*), cpartx(*), cparty(*), repc(*), xp(*), yp(*), up(*), vp(*)
      integer(c_int32_t), intent(in) :: nTOutBnd(*)
    end function
    integer(c_int) function wolfd2_b200_get_particles(ctx, xp, yp, up, vp, nTOutBnd) &
        bind(C, name='wolfd2_b200_get_particles')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: xp(*), yp(*), up(*), vp(*)
      integer(c_int32_t), intent(out) :: nTOutBnd(*)
    end function

    ! VelAvg / PTDAvg of the resident fields (set 0: u,v,p; set 1: uss,vss,pss) for output dumps, main.f:1053-1062
    integer(c_int) function wolfd2_b200_node_averages(ctx, set, util, vbar, pav, tav) &
        bind(C, name='wolfd2_b200_node_averages')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: set
      real(c_double), intent(out) :: util(*), vbar(*), pav(*), tav(*)   ! TAveraged of t / tss in tav
    end function

    ! several GPUs (one process each): slab layout, NCCL communicator, slab context
    integer(c_int) function wolfd2_b200_slab_layout(nx, ny, world, rank, lay) bind(C, name='wolfd2_b200_slab_layout')
      import :: c_int, c_int32_t
      integer(c_int32_t), value :: nx, ny, world, rank
      integer(c_int32_t), intent(out) :: lay(5)
    end function
    integer(c_int) function wolfd2_b200_set_device(device) bind(C, name='wolfd2_b200_set_device')
      import :: c_int, c_int32_t
      integer(c_int32_t), value :: device
    end function
    integer(c_int) function wolfd2_b200_comm_unique_id(id) bind(C, name='wolfd2_b200_comm_unique_id')
      import :: c_int, c_int8_t
      integer(c_int8_t), intent(out) :: id(128)
    end function
    integer(c_int) function wolfd2_b200_comm_init(rank, world, id) bind(C, name='wolfd2_b200_comm_init')
      import :: c_int, c_int32_t, c_int8_t
      integer(c_int32_t), value :: rank, world
      integer(c_int8_t), intent(in) :: id(128)
    end function
    integer(c_int) function wolfd2_b200_create_slab(ctx, par, reg, met, rank, world) bind(C, name='wolfd2_b200_create_slab')
      import :: c_int, c_int32_t, c_ptr, w2_params, w2_regions, w2_metrics
      type(c_ptr), intent(out) :: ctx
      type(w2_params),  intent(in) :: par
      type(w2_regions), intent(in) :: reg
      type(w2_metrics), intent(in) :: met
      integer(c_int32_t), value :: rank, world
    end function
    ! metric rows streamed in window by window (create with NULL metric pointers): which = 0-based position in w2_metrics
    integer(c_int) function wolfd2_b200_upload_metric_rows(ctx, which, jfirst, nrows, host) &
        bind(C, name='wolfd2_b200_upload_metric_rows')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: which, jfirst, nrows
      real(c_double), intent(in) :: host(*)
    end function
    ! a slab run gathered into a one-GPU context on rank 0 and compared bit for bit (verification)
    integer(c_int) function wolfd2_b200_gather_global(slab, global, what) bind(C, name='wolfd2_b200_gather_global')
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: slab, global
      integer(c_int32_t), value :: what
    end function
    integer(c_int) function wolfd2_b200_compare_global(slab, global, which, ndiff, maxabs) &
        bind(C, name='wolfd2_b200_compare_global')
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: slab, global
      integer(c_int32_t), value :: which
      integer(c_int64_t), intent(out) :: ndiff
      real(c_double), intent(out) :: maxabs
    end function

    ! time-series monitor points (SaveTimeSrs, file_manip.f:686-854; main.f:984-995): sampled inside wolfd2_b200_step
    integer(c_int) function wolfd2_b200_set_probes(ctx, npoints, iTS, jTS, freq) bind(C, name='wolfd2_b200_set_probes')
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: npoints, freq
      integer(c_int32_t), intent(in) :: iTS(*), jTS(*)
    end function
    integer(c_int) function wolfd2_b200_get_probe_records(ctx, maxrec, rec, steps, nrec) &
        bind(C, name='wolfd2_b200_get_probe_records')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: maxrec
      real(c_double), intent(out) :: rec(8, *)          ! rec(var, point + npoints*(record-1))
      integer(c_int32_t), intent(out) :: steps(*), nrec
    end function

    ! time averaging (-D_TIMEAVG_): op 0 begin (main.f:517-541), 1 / 2 accumulate pass 1 / 2 in every step (:1107-1208),
    ! 3 stop, 4 release; finish = :1239-1297; get: k = 0..18 (ubar vbar tbar pbar upb vpb tpb upupb vpvpb upvpb uptpb
    ! vptpb upxsb upysb vpxsb vpysb trbke dssrt dtdyb)
    integer(c_int) function wolfd2_b200_timeavg(ctx, op) bind(C, name='wolfd2_b200_timeavg')
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: op
    end function
    integer(c_int) function wolfd2_b200_timeavg_finish(ctx, npass, nts, uref, dlref) bind(C, name='wolfd2_b200_timeavg_finish')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: npass, nts
      real(c_double), value :: uref, dlref
    end function
    integer(c_int) function wolfd2_b200_timeavg_get(ctx, k, host) bind(C, name='wolfd2_b200_timeavg_get')
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: k
      real(c_double), intent(out) :: host(*)
    end function
    ! host waits issued by the last wolfd2_b200_step call (loop control lives on the device)
    integer(c_int) function wolfd2_b200_last_host_syncs(ctx, syncs) bind(C, name='wolfd2_b200_last_host_syncs')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int64_t), intent(out) :: syncs
    end function
  end interface
end module wolfd2_gpu
