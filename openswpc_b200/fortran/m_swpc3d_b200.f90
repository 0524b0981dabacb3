!! m_swpc3d_b200.f90 -- ISO_C_BINDING face of libswpc3d_b200.so (include/swpc3d_b200.h) for OpenSWPC's swpc_3d.
!!
!! SOURCE ONLY: the build image has no Fortran compiler, so this module is not compiled or tested there.  It shows the
!! binding a maintainer adds to src/swpc_3d/ so that the hot subroutines become one-line calls (INTEGRATION.md):
!!
!!   kernel__update_stress + absorb__update_stress -> swpc3d_update_stress     (m_kernel.f90:142, m_absorb.f90:60)
!!   source__stressglut(it)                        -> swpc3d_stressglut        (m_source.f90:776)
!!   global__comm_stress                           -> swpc3d_comm_stress       (m_global.f90:500)
!!   kernel__update_vel + absorb__update_vel       -> swpc3d_update_vel        (m_kernel.f90:75, m_absorb.f90:43)
!!   source__bodyforce(it)                         -> swpc3d_bodyforce         (m_source.f90:850)
!!   global__comm_vel                              -> swpc3d_comm_vel          (m_global.f90:391)
!!   wav__store(it) (velocity traces)              -> swpc3d_wav_store         (m_wav.f90:515-539)
!!   kernel__vmax                                  -> swpc3d_vmax              (m_kernel.f90:350)
!!   `!$acc enter data copyin(...)`                -> swpc3d_create / upload_medium / setup_pml|cerjan / set_sources /
!!                                                    set_stations            (main.f90:80-113)
module m_swpc3d_b200

    use iso_c_binding
    implicit none
    private

    integer(c_int32_t), parameter, public :: SWPC3D_ABC_PML = 1, SWPC3D_ABC_CERJAN = 2

    !! mirrors `swpc3d_grid` field by field
    type, bind(c), public :: swpc3d_grid
        integer(c_int32_t) :: nx, ny, nz
        integer(c_int32_t) :: nproc_x, nproc_y, myid
        integer(c_int32_t) :: ibeg, iend, jbeg, jend
        integer(c_int32_t) :: ipad, jpad, kpad
        integer(c_int32_t) :: ibeg_k, iend_k, jbeg_k, jend_k, kbeg_k, kend_k
        integer(c_int32_t) :: na, nm, abc_type, field_bytes, device, reserved
        real(c_double)     :: dx, dy, dz
        real(c_float)      :: dt, reserved_f
    end type swpc3d_grid

    public :: swpc3d_create, swpc3d_destroy, swpc3d_upload_medium, swpc3d_upload_fields, swpc3d_download_fields
    public :: swpc3d_setup_pml, swpc3d_setup_cerjan, swpc3d_set_sources, swpc3d_set_stations
    public :: swpc3d_update_stress, swpc3d_stressglut, swpc3d_comm_stress
    public :: swpc3d_update_vel, swpc3d_bodyforce, swpc3d_comm_vel
    public :: swpc3d_wav_store, swpc3d_step, swpc3d_sync, swpc3d_vmax, swpc3d_vmax_global, swpc3d_get_wav
    public :: swpc3d_nccl_unique_id, swpc3d_comm_init, swpc3d_last_error
    public :: swpc3d_set_wav_products, swpc3d_get_wav_product, swpc3d_set_option
    public :: swpc3d_snap_cfg, swpc3d_snap_setup, swpc3d_snap_step, swpc3d_snap_fetch, swpc3d_snap_fetch_max, swpc3d_reduce_sum
    public :: swpc3d_snap_fetch_begin, swpc3d_snap_fetch_end
    public :: swpc3d_set_green, swpc3d_green_store, swpc3d_green_source, swpc3d_get_green, swpc3d_advance
    public :: swpc3d_version, swpc3d_zero_state, swpc3d_run, swpc3d_timer_start, swpc3d_timer_stop, swpc3d_get_info, swpc3d_comm_local
    public :: swpc3d_check

    !! mirrors `swpc3d_snap_cfg` of include/swpc3d_b200.h: the integers snap__setup computes (m_snap.f90:116-154)
    type, bind(c) :: swpc3d_snap_cfg
        integer(c_int32_t) :: idec, jdec, kdec, ntdec_s
        integer(c_int32_t) :: nxs, nys, nzs
        integer(c_int32_t) :: is0, is1, js0, js1, ks0, ks1
        integer(c_int32_t) :: k0_xy, i0_yz, j0_xz
        integer(c_int32_t) :: sw(15)          !! xy_ps xy_v xy_u xz_ps xz_v xz_u yz_ps yz_v yz_u fs_ps fs_v fs_u ob_ps ob_v ob_u
        real(c_float) :: M0, UC
    end type swpc3d_snap_cfg

    interface

        function swpc3d_last_error() bind(c, name='swpc3d_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function

        integer(c_int) function swpc3d_create(g, ts, h) bind(c, name='swpc3d_create')
            import :: c_int, c_float, c_ptr, swpc3d_grid
            type(swpc3d_grid), intent(in) :: g
            real(c_float), intent(in) :: ts(*)          !! ts(1:nm), m_global.f90:41
            type(c_ptr), intent(out) :: h
        end function

        integer(c_int) function swpc3d_destroy(h) bind(c, name='swpc3d_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function

        !! arrays are passed exactly as allocated in m_medium.f90:441-452 (kbeg_m:kend_m, ibeg_m:iend_m, jbeg_m:jend_m)
        integer(c_int) function swpc3d_upload_medium(h, rho, lam, mu, taup, taus, kfs, kob, kfs_top, kfs_bot, kob_top, &
                                                     kob_bot, kbeg_a) bind(c, name='swpc3d_upload_medium')
            import :: c_int, c_float, c_int32_t, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: rho(*), lam(*), mu(*), taup(*), taus(*)
            integer(c_int32_t), intent(in) :: kfs(*), kob(*), kfs_top(*), kfs_bot(*), kob_top(*), kob_bot(*), kbeg_a(*)
        end function

        integer(c_int) function swpc3d_upload_fields(h, Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy) &
            bind(c, name='swpc3d_upload_fields')
            import :: c_int, c_ptr
            type(c_ptr), value :: h, Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy   !! c_loc(array) or c_null_ptr
        end function

        integer(c_int) function swpc3d_download_fields(h, Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy) &
            bind(c, name='swpc3d_download_fields')
            import :: c_int, c_ptr
            type(c_ptr), value :: h, Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy
        end function

        !! gxc(4,ibeg:iend) ... gze(4,kbeg:kend), m_absorb_p.f90:75-77
        integer(c_int) function swpc3d_setup_pml(h, gxc, gxe, gyc, gye, gzc, gze) bind(c, name='swpc3d_setup_pml')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: gxc(*), gxe(*), gyc(*), gye(*), gzc(*), gze(*)
        end function

        !! gx_c(ibeg_m:iend_m) ... gz_b(kbeg_m:kend_m), m_absorb_c.f90:45-47
        integer(c_int) function swpc3d_setup_cerjan(h, gx_c, gx_b, gy_c, gy_b, gz_c, gz_b) bind(c, name='swpc3d_setup_cerjan')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(in) :: gx_c(*), gx_b(*), gy_c(*), gy_b(*), gz_c(*), gz_b(*)
        end function

        !! mo..mxy are real(MP) = real(c_double) in the default build (m_source.f90:32-34); bf_mode: pass fx,fy,fz as mxx,myy,mzz
        integer(c_int) function swpc3d_set_sources(h, nsrc, isrc, jsrc, ksrc, mo, mxx, myy, mzz, myz, mxz, mxy, srcprm, &
                                                   stftype, bf_mode, tbeg) bind(c, name='swpc3d_set_sources')
            import :: c_int, c_int32_t, c_double, c_float, c_char, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: nsrc, bf_mode
            integer(c_int32_t), intent(in) :: isrc(*), jsrc(*), ksrc(*)
            real(c_double), intent(in) :: mo(*), mxx(*), myy(*), mzz(*), myz(*), mxz(*), mxy(*)
            real(c_float), intent(in) :: srcprm(*)
            character(kind=c_char), intent(in) :: stftype(*)   !! trim(stftype)//c_null_char
            real(c_float), value :: tbeg
        end function

        integer(c_int) function swpc3d_set_stations(h, nst, ist, jst, kst, ntdec_w, ntw, M0, UC) &
            bind(c, name='swpc3d_set_stations')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: nst, ntdec_w, ntw
            integer(c_int32_t), intent(in) :: ist(*), jst(*), kst(*)
            real(c_float), value :: M0, UC
        end function

        integer(c_int) function swpc3d_update_stress(h) bind(c, name='swpc3d_update_stress')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_stressglut(h, it) bind(c, name='swpc3d_stressglut')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_comm_stress(h) bind(c, name='swpc3d_comm_stress')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_update_vel(h) bind(c, name='swpc3d_update_vel')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_bodyforce(h, it) bind(c, name='swpc3d_bodyforce')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_comm_vel(h) bind(c, name='swpc3d_comm_vel')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_wav_store(h, it) bind(c, name='swpc3d_wav_store')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_step(h, it) bind(c, name='swpc3d_step')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_sync(h) bind(c, name='swpc3d_sync')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_vmax(h, vm) bind(c, name='swpc3d_vmax')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: vm(3)
        end function
        integer(c_int) function swpc3d_vmax_global(h, vm) bind(c, name='swpc3d_vmax_global')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: vm(3)
        end function
        integer(c_int) function swpc3d_get_wav(h, wav_vel) bind(c, name='swpc3d_get_wav')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: wav_vel(*)     !! wav_vel(ntw,3,nst), m_wav.f90:106
        end function
        integer(c_int) function swpc3d_nccl_unique_id(id) bind(c, name='swpc3d_nccl_unique_id')
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
        end function
        integer(c_int) function swpc3d_comm_init(h, id, nranks, rank) bind(c, name='swpc3d_comm_init')
            import :: c_int, c_int32_t, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int32_t), value :: nranks, rank
        end function

        !! wav__setup switches sw_wav_v/u/stress/strain (m_wav.f90:74-77) and the `update self(wav_*)` of wav__write (:672-675);
        !! which = 0 velocity (ntw,3,nst), 1 displacement (ntw,3,nst), 2 stress (ntw,6,nst), 3 strain (ntw,6,nst)
        integer(c_int) function swpc3d_set_wav_products(h, sw_v, sw_u, sw_stress, sw_strain) bind(c, name='swpc3d_set_wav_products')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: sw_v, sw_u, sw_stress, sw_strain
        end function
        integer(c_int) function swpc3d_get_wav_product(h, which, buf) bind(c, name='swpc3d_get_wav_product')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: which
            real(c_float), intent(out) :: buf(*)
        end function

        !! snap__write on the device (m_snap.f90:919-948) and the reductions onto the I/O rank (:1053-1066, :2295-2305)
        integer(c_int) function swpc3d_snap_setup(h, cfg) bind(c, name='swpc3d_snap_setup')
            import :: c_int, c_ptr, swpc3d_snap_cfg
            type(c_ptr), value :: h
            type(swpc3d_snap_cfg), intent(in) :: cfg
        end function
        integer(c_int) function swpc3d_snap_step(h, it) bind(c, name='swpc3d_snap_step')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_snap_fetch(h, product, root, rbuf) bind(c, name='swpc3d_snap_fetch')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: product, root      !! product = 3*section + type, 0-based, order of cfg%sw
            real(c_float), intent(out) :: rbuf(*)           !! (n1, n2, nvar) on the root
        end function
        integer(c_int) function swpc3d_snap_fetch_max(h, product, root, rbuf) bind(c, name='swpc3d_snap_fetch_max')
            import :: c_int, c_int32_t, c_float, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: product, root
            real(c_float), intent(out) :: rbuf(*)           !! (nxs, nys, 3): max-V, max-H, max-A
        end function
        !! the asynchronous pair: mpi_ireduce of the current record / mpi_wait before the next one (m_snap.f90:1057-1064)
        integer(c_int) function swpc3d_snap_fetch_begin(h, product, root, slot) bind(c, name='swpc3d_snap_fetch_begin')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: product, root, slot    !! slot = 0 or 1: which pinned host buffer receives the record
        end function
        integer(c_int) function swpc3d_snap_fetch_end(h, product, slot, rbuf) bind(c, name='swpc3d_snap_fetch_end')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: product, slot
            type(c_ptr), intent(out) :: rbuf                    !! c_f_pointer(rbuf, buf, [n1, n2, nvar]) on the root; c_null_ptr elsewhere
        end function
        integer(c_int) function swpc3d_reduce_sum(h, buf, n, root) bind(c, name='swpc3d_reduce_sum')
            import :: c_int, c_int32_t, c_int64_t, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(inout) :: buf(*)
            integer(c_int64_t), value :: n
            integer(c_int32_t), value :: root
        end function

        !! Green's-function mode (m_green.f90): device copy-in of green__setup (:351), green__store, green__source, update self(gf)
        integer(c_int) function swpc3d_set_green(h, ng, ig, jg, kg, bforce, is_src, isrc, jsrc, ksrc, fx1, fy1, fz1, trise, &
                                                 stftype, ntdec_w, ntw, tbeg) bind(c, name='swpc3d_set_green')
            import :: c_int, c_int32_t, c_float, c_char, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: ng, bforce, is_src, isrc, jsrc, ksrc, ntdec_w, ntw
            integer(c_int32_t), intent(in) :: ig(*), jg(*), kg(*)
            real(c_float), value :: fx1, fy1, fz1, trise, tbeg
            character(kind=c_char), intent(in) :: stftype(*)
        end function
        integer(c_int) function swpc3d_advance(h, it) bind(c, name='swpc3d_advance')   !! main.f90:126-138 (stress .. comm_vel)
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_green_store(h, it) bind(c, name='swpc3d_green_store')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_green_source(h, it) bind(c, name='swpc3d_green_source')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it
        end function
        integer(c_int) function swpc3d_get_green(h, gf) bind(c, name='swpc3d_get_green')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: gf(*)             !! gf(ntw, ncmp*ng) as m_green.f90:325
        end function

        !! tuning / modes: key is a null-terminated name ("pw_mode", "tma", "jlen", "overlap", ...)
        integer(c_int) function swpc3d_set_option(h, key, value) bind(c, name='swpc3d_set_option')
            import :: c_int, c_int32_t, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: key(*)
            integer(c_int32_t), value :: value
        end function

        !! ---- the rest of the C ABI: run loop, stopwatch (m_pwatch replacement), introspection, test hooks
        function swpc3d_version() bind(c, name='swpc3d_version') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        integer(c_int) function swpc3d_zero_state(h) bind(c, name='swpc3d_zero_state')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        !! it0 .. it1 iterations of swpc3d_step without a host round trip
        integer(c_int) function swpc3d_run(h, it0, it1) bind(c, name='swpc3d_run')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), value :: h
            integer(c_int32_t), value :: it0, it1
        end function
        integer(c_int) function swpc3d_timer_start(h) bind(c, name='swpc3d_timer_start')
            import :: c_int, c_ptr
            type(c_ptr), value :: h
        end function
        integer(c_int) function swpc3d_timer_stop(h, ms) bind(c, name='swpc3d_timer_stop')
            import :: c_int, c_float, c_ptr
            type(c_ptr), value :: h
            real(c_float), intent(out) :: ms
        end function
        integer(c_int) function swpc3d_get_info(h, key, value) bind(c, name='swpc3d_get_info')
            import :: c_int, c_double, c_char, c_ptr
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: key(*)      !! "launches", "ms_stress", "ms_vel", ... (NUL-terminated)
            real(c_double), intent(out) :: value
        end function
        !! several ranks living on one GPU (tests): handles(n) ordered by myid; which = 0 stress, 1 velocity
        integer(c_int) function swpc3d_comm_local(handles, n, which) bind(c, name='swpc3d_comm_local')
            import :: c_int, c_int32_t, c_ptr
            type(c_ptr), intent(in) :: handles(*)
            integer(c_int32_t), value :: n, which
        end function

    end interface

contains

    !! non-zero return -> the reference's convention: message + stop (m_debug.f90:206-221)
    subroutine swpc3d_check(ierr)
        use iso_fortran_env, only: error_unit
        integer(c_int), intent(in) :: ierr
        character(kind=c_char), pointer :: msg(:)
        integer :: n
        if (ierr == 0) return
        call c_f_pointer(swpc3d_last_error(), msg, [512])
        n = 1
        do while (n < 512 .and. msg(n) /= c_null_char); n = n + 1; end do
        write (error_unit, '(A,512A1)') '[swpc3d_b200] ', msg(1:n - 1)
        stop 1
    end subroutine swpc3d_check

end module m_swpc3d_b200
