!! m_swpcpsv_b200.f90 -- ISO_C_BINDING face of the swpc_psv entry points of libswpc3d_b200.so (include/swpcpsv_b200.h).
!!
!! SOURCE ONLY (no Fortran compiler in the build image), like m_swpc3d_b200.f90.  The hot subroutines of src/swpc_psv become
!! one-line calls (INTEGRATION.md 3b):
!!
!!   kernel__update_stress + absorb__update_stress           -> swpcpsv_update_stress   (m_kernel.f90:142, m_absorb.f90:60)
!!   source__stressglut(it)                                  -> swpcpsv_stressglut      (m_source.f90:550)
!!   global__comm_stress                                     -> swpcpsv_comm_stress     (m_global.f90:366)
!!   kernel__update_vel + source__bodyforce + absorb__update_vel -> swpcpsv_update_vel(it) (main.f90:108-110, the P-SV order)
!!   global__comm_vel                                        -> swpcpsv_comm_vel        (m_global.f90:312)
!!   wav__store(it)                                          -> swpcpsv_wav_store       (m_wav.f90:143-306)
!!   snap__write(it) (device part) / mpi_reduce of a slice   -> swpcpsv_snap_step / swpcpsv_snap_fetch (m_snap.f90:435-650)
!!   kernel__vmax                                            -> swpcpsv_vmax            (m_kernel.f90:313)
!!   `!$acc enter data copyin(...)` of main.f90:80-93        -> swpcpsv_create / upload_medium / setup_pml|cerjan /
!!                                                              set_sources / set_stations
module m_swpcpsv_b200

    use iso_c_binding
    implicit none
    private

    integer(c_int32_t), parameter, public :: SWPCPSV_ABC_PML = 1, SWPCPSV_ABC_CERJAN = 2

    !! mirrors `swpcpsv_grid` field by field
    type, bind(c), public :: swpcpsv_grid
        integer(c_int32_t) :: nx, nz
        integer(c_int32_t) :: nproc_x, myid
        integer(c_int32_t) :: ibeg, iend
        integer(c_int32_t) :: ipad, kpad
        integer(c_int32_t) :: ibeg_k, iend_k, kend_k
        integer(c_int32_t) :: na, nm, abc_type, field_bytes, device
        real(c_double)     :: dx, dz
        real(c_float)      :: dt, reserved_f
    end type swpcpsv_grid

    !! mirrors `swpcpsv_snap_cfg`: the integers snap__setup computes (m_snap.f90:94-114) and the output scaling
    type, bind(c), public :: swpcpsv_snap_cfg
        integer(c_int32_t) :: idec, kdec, ntdec_s, nxs, nzs, is0, is1, ks0, ks1, sw_ps, sw_v, sw_u
        real(c_float)      :: M0, UC
    end type swpcpsv_snap_cfg

    public :: swpcpsv_last_error, swpcpsv_create, swpcpsv_destroy, swpcpsv_upload_medium, swpcpsv_upload_fields, swpcpsv_download_fields
    public :: swpcpsv_setup_pml, swpcpsv_setup_cerjan, swpcpsv_set_sources, swpcpsv_set_stations, swpcpsv_get_wav
    public :: swpcpsv_update_stress, swpcpsv_stressglut, swpcpsv_comm_stress, swpcpsv_update_vel, swpcpsv_comm_vel
    public :: swpcpsv_wav_store, swpcpsv_step, swpcpsv_sync, swpcpsv_vmax, swpcpsv_vmax_global
    public :: swpcpsv_snap_setup, swpcpsv_snap_step, swpcpsv_snap_fetch, swpcpsv_reduce_sum
    public :: swpcpsv_version, swpcpsv_zero_state, swpcpsv_run, swpcpsv_timer_start, swpcpsv_timer_stop, swpcpsv_get_info, swpcpsv_comm_local, swpcpsv_download_memvars
    public :: swpcpsv_nccl_unique_id, swpcpsv_comm_init, swpcpsv_set_option, swpcpsv_check

    interface

        function swpcpsv_last_error() bind(c, name='swpcpsv_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function

        integer(c_int) function swpcpsv_create(g, ts, h) bind(c, name='swpcpsv_create')
            import :: c_int, c_float, c_ptr, swpcpsv_grid
            type(swpcpsv_grid), intent(in) :: g
            real(c_float), intent(in) :: ts(*)                  !! ts(1:nm), visco_set_relaxtime
            type(c_ptr), intent(out) :: h
        end function
        integer(c_int) function swpcpsv_destroy(h) bind(c, name='swpcpsv_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function

        !! arrays exactly as allocated in m_medium / m_global: (kbeg_m:kend_m, ibeg_m:iend_m) and (ibeg_m:iend_m)
        integer(c_int) function swpcpsv_upload_medium(h, rho, lam, mu, taup, taus, kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot, kbeg_a) &
            bind(c, name='swpcpsv_upload_medium')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: rho(*), lam(*), mu(*), taup(*), taus(*)
            integer(c_int32_t), intent(in) :: kfs(*), kob(*), kfs_top(*), kfs_bot(*), kob_top(*), kob_bot(*), kbeg_a(*)
        end function
        !! Vx Vz Sxx Szz Sxz as real(MP) arrays (c_loc of each); c_null_ptr leaves a field untouched (plane-wave mode uploads all)
        integer(c_int) function swpcpsv_upload_fields(h, Vx, Vz, Sxx, Szz, Sxz) bind(c, name='swpcpsv_upload_fields')
            import :: c_int, c_ptr
            type(c_ptr), value :: h, Vx, Vz, Sxx, Szz, Sxz
        end function
        integer(c_int) function swpcpsv_download_fields(h, Vx, Vz, Sxx, Szz, Sxz) bind(c, name='swpcpsv_download_fields')
            import :: c_int, c_ptr
            type(c_ptr), value :: h, Vx, Vz, Sxx, Szz, Sxz
        end function

        integer(c_int) function swpcpsv_setup_pml(h, gxc, gxe, gzc, gze) bind(c, name='swpcpsv_setup_pml')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: gxc(4, *), gxe(4, *), gzc(4, *), gze(4, *)   !! m_absorb_p.f90:57-101
        end function
        integer(c_int) function swpcpsv_setup_cerjan(h, gx_c, gx_b, gz_c, gz_b) bind(c, name='swpcpsv_setup_cerjan')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: gx_c(*), gx_b(*), gz_c(*), gz_b(*)            !! m_absorb_c.f90:28-96
        end function

        !! moment mode: mo (already / M0), mxx, mzz, mxz; body-force mode (bf_mode /= 0): fx, fz in mxx, mzz
        integer(c_int) function swpcpsv_set_sources(h, nsrc, isrc, ksrc, mo, mxx, mzz, mxz, srcprm, stftype, bf_mode, tbeg) &
            bind(c, name='swpcpsv_set_sources')
            import :: c_int, c_int32_t, c_double, c_float, c_char, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: nsrc, bf_mode
            integer(c_int32_t), intent(in) :: isrc(*), ksrc(*)
            real(c_double), intent(in) :: mo(*), mxx(*), mzz(*), mxz(*)
            real(c_float), intent(in) :: srcprm(2, *)
            character(kind=c_char), intent(in) :: stftype(*)
            real(c_float), value :: tbeg
        end function
        integer(c_int) function swpcpsv_set_stations(h, nst, ist, kst, ntdec_w, ntw, M0, UC, sw_v, sw_u, sw_stress, sw_strain) &
            bind(c, name='swpcpsv_set_stations')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: nst, ntdec_w, ntw, sw_v, sw_u, sw_stress, sw_strain
            integer(c_int32_t), intent(in) :: ist(*), kst(*)
            real(c_float), value :: M0, UC
        end function
        !! which: 0 velocity (ntw,2,nst), 1 displacement (ntw,2,nst), 2 stress (ntw,3,nst), 3 strain (ntw,3,nst)
        integer(c_int) function swpcpsv_get_wav(h, which, wav) bind(c, name='swpcpsv_get_wav')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: which
            real(c_float), intent(out) :: wav(*)
        end function

        integer(c_int) function swpcpsv_update_stress(h) bind(c, name='swpcpsv_update_stress')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpcpsv_stressglut(h, it) bind(c, name='swpcpsv_stressglut')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpcpsv_comm_stress(h) bind(c, name='swpcpsv_comm_stress')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpcpsv_update_vel(h, it) bind(c, name='swpcpsv_update_vel')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpcpsv_comm_vel(h) bind(c, name='swpcpsv_comm_vel')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpcpsv_wav_store(h, it) bind(c, name='swpcpsv_wav_store')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpcpsv_step(h, it) bind(c, name='swpcpsv_step')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpcpsv_sync(h) bind(c, name='swpcpsv_sync')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpcpsv_vmax(h, vm) bind(c, name='swpcpsv_vmax')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: vm(2)
        end function
        integer(c_int) function swpcpsv_vmax_global(h, vm) bind(c, name='swpcpsv_vmax_global')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: vm(2)
        end function

        integer(c_int) function swpcpsv_snap_setup(h, cfg) bind(c, name='swpcpsv_snap_setup')
            import :: c_int, c_ptr, swpcpsv_snap_cfg
            type(c_ptr), value :: h
            type(swpcpsv_snap_cfg), intent(in) :: cfg
        end function
        integer(c_int) function swpcpsv_snap_step(h, it) bind(c, name='swpcpsv_snap_step')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpcpsv_snap_fetch(h, product, root, rbuf) bind(c, name='swpcpsv_snap_fetch')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: product, root        !! 0 ps, 1 v, 2 u
            real(c_float), intent(out) :: rbuf(*)             !! (nxs, nzs, 2) on the root
        end function
        integer(c_int) function swpcpsv_reduce_sum(h, buf, n, root) bind(c, name='swpcpsv_reduce_sum')
            import :: c_int, c_int32_t, c_int64_t, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(inout) :: buf(*)
            integer(c_int64_t), value :: n
            integer(c_int32_t), value :: root
        end function

        integer(c_int) function swpcpsv_nccl_unique_id(id) bind(c, name='swpcpsv_nccl_unique_id')
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
        end function
        integer(c_int) function swpcpsv_comm_init(h, id, nranks, rank) bind(c, name='swpcpsv_comm_init')
            import :: c_int, c_int32_t, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int32_t), value :: nranks, rank
        end function

        integer(c_int) function swpcpsv_set_option(h, key, value) bind(c, name='swpcpsv_set_option')
            import :: c_int, c_int32_t, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: key(*)      !! "pw_mode", "tk", "ilen", "pf", ...
            integer(c_int32_t), value :: value
        end function

        !! ---- the rest of the C ABI: run loop, stopwatch (m_pwatch replacement), introspection, test hooks
        function swpcpsv_version() bind(c, name='swpcpsv_version') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        integer(c_int) function swpcpsv_zero_state(h) bind(c, name='swpcpsv_zero_state')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        !! it0 .. it1 iterations of swpcpsv_step without a host round trip
        integer(c_int) function swpcpsv_run(h, it0, it1) bind(c, name='swpcpsv_run')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it0, it1
        end function
        integer(c_int) function swpcpsv_timer_start(h) bind(c, name='swpcpsv_timer_start')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpcpsv_timer_stop(h, ms) bind(c, name='swpcpsv_timer_stop')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: ms
        end function
        integer(c_int) function swpcpsv_get_info(h, key, value) bind(c, name='swpcpsv_get_info')
            import :: c_int, c_double, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: key(*)      !! "launches", "ms_stress", "ms_vel", ... (NUL-terminated)
            real(c_double), intent(out) :: value
        end function
        !! several ranks living on one GPU (tests): handles(n) ordered by myid; which = 0 stress, 1 velocity
        integer(c_int) function swpcpsv_comm_local(handles, n, which) bind(c, name='swpcpsv_comm_local')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), intent(in) :: handles(*)
            integer(c_int32_t), value :: n, which
        end function
        !! memory variables in the reference layout (m, k, i) over the memory box (m_kernel.f90:337-339); c_null_ptr skips one
        integer(c_int) function swpcpsv_download_memvars(h, Rxx, Rzz, Rxz) bind(c, name='swpcpsv_download_memvars')
            import :: c_int, c_ptr
            type(c_ptr), value :: h, Rxx, Rzz, Rxz
        end function

    end interface

contains

    !! non-zero return -> the reference's convention: message + stop (m_debug.f90:206-221)
    subroutine swpcpsv_check(ierr)
        use iso_fortran_env, only: error_unit
        integer(c_int), intent(in) :: ierr
        character(kind=c_char), pointer :: msg(:)
        integer :: n
        if (ierr == 0) return
        call c_f_pointer(swpcpsv_last_error(), msg, [512])
        n = 1
        do while (n < 512 .and. msg(n) /= c_null_char); n = n + 1; end do
        write (error_unit, '(A,512A1)') '[swpcpsv_b200] ', msg(1:n - 1)
        stop 1
    end subroutine swpcpsv_check

end module m_swpcpsv_b200
