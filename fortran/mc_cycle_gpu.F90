! -*- mode: F90 -*-
!=============================================================================!
! Coarse GPU path for mc_moves.F90 of keb721/mc_water_ls_mw.                  !
!                                                                             !
! This file is a PATCH FRAGMENT, not a module: the three routines below are   !
! added to the `contains` section of module mc_moves (they use its private    !
! variables), and two call sites change:                                      !
!                                                                             !
!  (1) end of mc_init (mc_moves.F90:862, after ls_mu is first computed):      !
!          call mc_gpu_init()                                                 !
!  (2) mc_cycle, mc_moves.F90:212-255 (list refresh + "do imove = 1,nwater"   !
!      loop + average_energy accumulation) are replaced by                    !
!          call mc_gpu_cycle()                                                !
!      and the comms_allreduce_* calls at :264-268 by                         !
!          call mwgpu_check(mwgpu_comms_allreduce_bins(gpu_ctx),'mc_cycle')   !
!          call mc_gpu_pull_bins()                                            !
!  (3) the periodic bookkeeping keeps its log lines and files on the host but  !
!      its STATE effects happen on the device -- there is no host-side path   !
!      that modifies the state and pushes it back (the device owns step       !
!      sizes, counters, running averages, energies, lists, *_last_sync bases):!
!        mc_monitor_stats               -> call mc_gpu_monitor()      (below) !
!        mc_check_chain_synchronisation -> call mc_gpu_chain_sync()   (below) !
!        mc_check_flatness              -> call mc_gpu_check_flatness()       !
!        mc_compute_deltaG_from_hist    -> call mc_gpu_deltaG()               !
!      mc_checkpoint_write / dcd output only READ:  call mc_gpu_pull()  first. !
!      mc_gpu_push() is for start-up and restart only (it rebuilds the lists  !
!      and the energies as energy_init does).                                 !
!                                                                             !
! Everything else in mc_moves.F90 / main.f90 / io.f90 is untouched: program   !
! entry, input decks and output formats stay as they are.                     !
!                                                                             !
! NOT COMPILED in this repository (no Fortran compiler in the build image).   !
!=============================================================================!

  subroutine mc_gpu_init()
    !--------------------------------------------------------------------------!
    ! Hands the run parameters, bin grid inputs, weights and the random stream !
    ! to the device (replaces nothing; mirrors mc_init :504-877 on the device).!
    !--------------------------------------------------------------------------!
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx
    use comms,      only : myrank,size
    use model,      only : ls
    use userparams
    implicit none
    type(mwgpu_mc_params) :: p

    p%temperature = temperature          ; p%pressure = pressure          ! internal units (io.f90:165)
    p%npt = merge(1,0,mc_ensemble=='npt')
    p%mc_max_trans = mc_max_trans        ; p%mc_dv_max = mc_dv_max        ! Bohr (io.f90:185-186)
    p%mc_target_ratio = mc_target_ratio
    p%wl_factor = orig_wl_factor         ; p%wl_swetnam = merge(1,0,wl_swetnam)   ! userparams.f90:37 (wl_schedule travels in mwgpu_flat_params)
    p%wl_alpha = wl_alpha
    p%eta_interp = merge(1,0,eta_interp) ; p%samplerun = merge(1,0,samplerun)
    p%leshift = merge(1,0,leshift)       ; p%nbins = nbins
    p%mu_min = mu_min                    ; p%mu_max = mu_max
    p%allow_switch = merge(1,0,allow_switch) ; p%allow_vol = merge(1,0,allow_vol)
    p%allow_trans = merge(1,0,allow_trans)
    p%mc_trans_prob = mc_trans_prob      ; p%mc_vol_prob = mc_vol_prob
    p%mc_switch_prob = mc_switch_prob    ; p%mc_always_switch = merge(1,0,mc_always_switch)
    p%list_update_int = list_update_int  ; p%eq_mc_cycles = eq_mc_cycles
    p%max_mc_cycles = max_mc_cycles      ; p%eq_adjust_mc = merge(1,0,eq_adjust_mc)
    p%monitor_int = monitor_int          ; p%dd = merge(1,0,parallel_strategy=='dd')
    p%window_overlap = window_overlap
    p%input_ref_enthalpy = input_ref_enthalpy
    p%ls = ls

    ! one walker per rank, rank = myrank of size ranks (windows in dd mode, log_unbiased_norm)
    call mwgpu_check(mwgpu_mc_init(gpu_ctx,p,int(myrank,c_int),int(max(size,1),c_int), &
                                   weight,int(nbins,c_int),wl_factor),'mc_gpu_init')
    ! host-side state that mc_init may have restored from a checkpoint
    call mc_gpu_push()
    ! random stream: the compiler's random_number is unpinned (random.f90:62-63); the device
    ! uses Philox-4x32-10 keyed by (seed, rank); main.f90:79-81 burns 1 000 000 draws first.
    call mwgpu_check(mwgpu_mc_set_rng_philox(gpu_ctx,20141211_c_int64_t,int(myrank,c_int32_t), &
                                             1000000_c_int64_t),'mc_gpu_init')
  end subroutine mc_gpu_init

  subroutine mc_gpu_cycle()
    !--------------------------------------------------------------------------!
    ! mc_moves.F90:212-255 for one cycle on the device, then the scalars the    !
    ! rest of mc_cycle reads.                                                   !
    !--------------------------------------------------------------------------!
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx,model_energy
    use model,      only : volume,ls
    use userparams, only : num_lattices,mc_max_trans,mc_dv_max,wl_factor
    implicit none
    type(mwgpu_walker_state) :: st
    call mwgpu_check(mwgpu_mc_run(gpu_ctx,1_c_int),'mc_gpu_cycle')
    call mwgpu_check(mwgpu_mc_get_state(gpu_ctx,0_c_int,st),'mc_gpu_cycle')
    model_energy(1:num_lattices) = st%model_energy(1:num_lattices)
    volume(1:num_lattices)       = st%volume(1:num_lattices)
    average_energy               = st%average_energy
    ls_mu = st%ls_mu ; ls = st%ls
    mc_accepted_rsteps  = st%accepted(1)  ; mc_accepted_vsteps  = st%accepted(2)  ; mc_accepted_swtch  = st%accepted(3)
    mc_attempted_rsteps = st%attempted(1) ; mc_attempted_vsteps = st%attempted(2) ; mc_attempted_swtch = st%attempted(3)
    min_dmu = st%min_dmu ; max_dmu = st%max_dmu
    wl_factor = st%wl_factor
    walker_in_window = (st%walker_in_window/=0)
  end subroutine mc_gpu_cycle

  subroutine mc_gpu_pull_bins()
    use iso_c_binding
    use mwgpu
    use energy, only : gpu_ctx
    implicit none
    call mwgpu_check(mwgpu_mc_get_bins(gpu_ctx,0_c_int,weight,histogram,unbiased_hist),'mc_gpu_pull_bins')
  end subroutine mc_gpu_pull_bins

  subroutine mc_gpu_pull()
    ! device -> host before host-side bookkeeping (monitor, flatness, chain sync, checkpoint, dcd)
    use iso_c_binding
    use mwgpu
    use energy, only : gpu_ctx,energy_pull_from_device
    implicit none
    call energy_pull_from_device()
    call mc_gpu_pull_bins()
    call mwgpu_check(mwgpu_mc_get_translations(gpu_ctx,0_c_int,mc_translations),'mc_gpu_pull')
  end subroutine mc_gpu_pull

  subroutine mc_gpu_push()
    ! host -> device at start-up / after a checkpoint load ONLY: re-initialises lists and energies like energy_init.
    ! The periodic bookkeeping never goes through here (see (3) above).
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx,energy_push_to_device
    use model,      only : ls
    use userparams, only : wl_factor
    implicit none
    call energy_push_to_device()
    call mwgpu_check(mwgpu_energy_init(gpu_ctx),'mc_gpu_push')          ! lists, energies, ls_mu (:842-862)
    call mwgpu_check(mwgpu_mc_set_bins(gpu_ctx,0_c_int,weight,histogram,unbiased_hist),'mc_gpu_push')
    call mwgpu_check(mwgpu_mc_set_wl_factor(gpu_ctx,0_c_int,wl_factor,merge(1_c_int,0_c_int,wl_invt_active)),'mc_gpu_push')
    call mwgpu_check(mwgpu_mc_set_active_lattice(gpu_ctx,0_c_int,int(ls,c_int)),'mc_gpu_push')
  end subroutine mc_gpu_push

  subroutine mc_gpu_monitor()
    ! mc_monitor_stats (mc_moves.F90:1692-1850): the host prints what it always printed from the pulled
    ! counters; step-size adjustment (:1722-1732), energy re-synchronisation (:1786-1792), counter reset
    ! (:1797-1810) and the all-reduce of histogram / weights / unbiased histogram (:1813-1821) run on the device
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx,model_energy
    use userparams, only : num_lattices,parallel_strategy,mc_max_trans,mc_dv_max
    implicit none
    type(mwgpu_walker_state) :: st
    call mc_gpu_pull()                    ! acceptance counters, mc_translations, average_energy for the log lines
    ! ... the unchanged write statements of :1734-1780 (ratios, attempts per molecule, mu span, energies) ...
    call mwgpu_check(mwgpu_mc_monitor(gpu_ctx),'mc_gpu_monitor')
    if (num_lattices==2 .and. parallel_strategy=='mw') then
       call mwgpu_check(mwgpu_comms_allreduce_bins(gpu_ctx),'mc_gpu_monitor')
       call mc_gpu_pull_bins()            ! for the histogram.dat / eta_weights.dat dumps of :1827-1847
    end if
    call mwgpu_check(mwgpu_mc_get_state(gpu_ctx,0_c_int,st),'mc_gpu_monitor')
    mc_max_trans = st%mc_max_trans ; mc_dv_max = st%mc_dv_max          ! tuned on the device during equilibration
    model_energy(1:num_lattices) = st%model_energy(1:num_lattices)     ! "Checking accumulated energies" (:1781-1792)
  end subroutine mc_gpu_monitor

  subroutine mc_gpu_chain_sync()
    ! mc_check_chain_synchronisation (mc_moves.F90:2217-2416) on the device; the report lines of :2404-2412 read
    ! the energies before / after from two mwgpu_mc_get_state calls
    use iso_c_binding
    use mwgpu
    use energy, only : gpu_ctx
    implicit none
    call mwgpu_check(mwgpu_mc_chain_sync(gpu_ctx),'mc_gpu_chain_sync')
  end subroutine mc_gpu_chain_sync

!=============================================================================!
! Device-side variants of the periodic bookkeeping (no pull / push round trip) !
!                                                                             !
!  (4) mc_check_flatness (mc_moves.F90:1936-2185): its body becomes            !
!          call mc_gpu_check_flatness()                                        !
!      which keeps the log lines and the wlf.dat / eta_weights.dat_* /         !
!      histogram.dat_* dumps, fed from the returned report and the pulled bins.!
!  (5) mc_compute_deltaG_from_hist (:2498-2621): the all-reduce / join, the    !
!      normalisation and the log(pA/pB) come from mwgpu_mc_deltag_from_hist;   !
!      the log lines and unbiased_histogram_<cycle>.dat stay.                  !
!  (6) main.f90:200-223: with  call mwgpu_mc_set_therm(gpu_ctx,file_output_int,!
!      capacity)  at start-up the kernel records the values of every therm row !
!      itself, so mc_gpu_cycle may advance many cycles per call;               !
!      mc_gpu_write_therm() drains and prints them with the unchanged formats. !
!  (7) restart: after mc_checkpoint_load has read the records (:449-492) the   !
!      refresh of mc_init :842-862 is  call mc_gpu_restore().                  !
!=============================================================================!

  subroutine mc_gpu_check_flatness()
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx
    use comms,      only : myrank
    use io,         only : glog
    use userparams, only : wl_schedule,wl_minhist,wl_flattol,wl_useinvt,wl_factor
    implicit none
    type(mwgpu_flat_params) :: fp
    type(mwgpu_flat_report) :: rep
    fp%wl_schedule = wl_schedule ; fp%wl_minhist = wl_minhist
    fp%wl_flattol  = wl_flattol  ; fp%wl_useinvt = merge(1,0,wl_useinvt)
    call mwgpu_check(mwgpu_mc_check_flatness(gpu_ctx,fp,rep),'mc_gpu_check_flatness')
    if (rep%checked==0) return                                  ! :1961
    wl_factor = rep%wl_factor
    wl_invt_active = wl_invt_active.or.(rep%invt_switched/=0)
    if (rep%flat/=0) firstcycle = .false.
    if (rep%hist_reset/=0) return                               ! :1973-1980
    if (myrank==0) then                                         ! :1993-2000
       write(glog,'("# Checking flatness of histogram at cycle ",I10,"           #")')mc_cycle_num
       write(glog,'("# Most  populated histogram bin = ",F10.4," % of mean         #")')rep%max_pct
       write(glog,'("# Least populated histogram bin = ",F10.4," % of mean         #")')rep%min_pct
    end if
    if ((rep%flat/=0).and.(myrank==0)) then
       call mc_gpu_pull_bins()                                  ! weights (shifted) for eta_weights.dat_<f>
       ! ... the unchanged file dumps of :2067-2101 ...
    end if
  end subroutine mc_gpu_check_flatness

  subroutine mc_gpu_deltaG(deltaG,normP)
    use iso_c_binding
    use mwgpu
    use energy, only : gpu_ctx
    implicit none
    real(kind=dp),intent(out) :: deltaG
    real(kind=dp),intent(out) :: normP(:)
    call mwgpu_check(mwgpu_mc_deltag_from_hist(gpu_ctx,deltaG,normP),'mc_gpu_deltaG')
  end subroutine mc_gpu_deltaG

  subroutine mc_gpu_write_therm(mytherm)
    ! main.f90:200-223 from the rows the kernel recorded since the last call
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx
    use constants,  only : hart_to_ev,bohr_to_ang,water_mass,aud_to_kgm3
    use userparams, only : num_lattices,samplerun,wl_factor,nwater
    use util,       only : util_hmatrix_to_abc
    implicit none
    integer,intent(in) :: mytherm
    integer,parameter  :: maxrows = 64
    type(mwgpu_therm_row) :: rows(maxrows)
    integer(c_int) :: n,ndrop
    integer :: i,l
    real(kind=dp) :: a,b,c,al,be,ga,h1(3,3)
    call mwgpu_check(mwgpu_mc_get_therm(gpu_ctx,0_c_int,rows,int(maxrows,c_int),n,ndrop),'mc_gpu_write_therm')
    if (ndrop/=0) stop 'therm ring overflow: drain more often or raise its capacity'
    do i = 1,n
       l = int(rows(i)%ls)
       if (num_lattices==1) then
          h1 = reshape(rows(i)%hmatrix1,(/3,3/))
          call util_hmatrix_to_abc(h1,a,b,c,al,be,ga)
          write(mytherm,'(I8,E15.6,5x,F15.6,6F15.6)')int(rows(i)%icyc),rows(i)%model_energy(1)*hart_to_ev, &
               rows(i)%volume(1)*bohr_to_ang**3,a*bohr_to_ang,b*bohr_to_ang,c*bohr_to_ang,al,be,ga
       else if ((wl_factor<tiny(1.0_dp)).or.samplerun) then
          write(mytherm,'(I8,E15.6,5x,3F15.6,1x,I1)')int(rows(i)%icyc),rows(i)%model_energy(l)*hart_to_ev, &
               rows(i)%ls_mu,rows(i)%volume*bohr_to_ang**3,l
       else
          write(mytherm,'(I8,E15.6,5x,2F15.6,1x,I1)')int(rows(i)%icyc),rows(i)%model_energy(l)*hart_to_ev, &
               rows(i)%ls_mu,real(nwater,kind=dp)*water_mass/rows(i)%volume(l)*aud_to_kgm3,l
       end if
    end do
  end subroutine mc_gpu_write_therm

  subroutine mc_gpu_restore()
    ! after mc_checkpoint_load (:403-501): replaces the host-side refresh of mc_init :842-862
    use iso_c_binding
    use mwgpu
    use energy,     only : gpu_ctx
    use model,      only : hmatrix,ljr,ref_ljr,ls
    use userparams, only : mc_max_trans,mc_dv_max,wl_factor
    implicit none
    call mwgpu_check(mwgpu_mc_restore(gpu_ctx,0_c_int,int(mc_cycle_num,c_int),mc_max_trans,mc_dv_max,wl_factor, &
         merge(1_c_int,0_c_int,wl_invt_active),int(ls,c_int),histogram,weight,unbiased_hist, &
         hmatrix,ref_ljr,ljr),'mc_gpu_restore')
  end subroutine mc_gpu_restore
