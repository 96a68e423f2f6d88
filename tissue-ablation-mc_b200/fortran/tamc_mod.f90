module tamc_mod
!
!  iso_c_binding interface to libtamc.so (include/tamc.h): the GPU photon transport that takes the
!  place of the per-rank photon loop and of the jmean reduction in mcpolar.f90 (lines 151-173 of the
!  reference driver).  Shipped as source: this image has no Fortran compiler, so the module is not
!  built or tested here (see INTEGRATION.md); the same entry points are exercised through the C++
!  driver shim (tissue-ablation-mc_b200/driver) and the Python ctypes binding.
!
!  The reference is compiled with -freal-4-real-8 (src/Makefile:3), so its default `real` arrays
!  (rhokap, jmean, jmeanGLOBAL in iarray.f90) are c_double and can be passed without a copy.
!
    use, intrinsic :: iso_c_binding

    implicit none

    integer(c_int), parameter :: TAMC_OK = 0
    integer(c_int), parameter :: TAMC_SCATTER = 1, TAMC_FRESNEL = 2, TAMC_PERIODIC = 4

    !  mirrors tamc_stats (include/tamc.h)
    type, bind(C) :: tamc_stats
        integer(c_int64_t) :: packets, voxel_steps, scatters, absorbed
        integer(c_int64_t) :: exits(6)
        real(c_double)     :: zero_ms, kernel_ms, allreduce_ms, h2d_ms, d2h_ms
        integer(c_int64_t) :: gpu_launches, specular, internal_reflections
    end type tamc_stats

    !  mirrors tamc_heat_params (include/tamc.h): the Heat-module inputs of res/input.params
    type, bind(C) :: tamc_heat_params
        real(c_double)     :: power, energyPerPixel, total_time, repetitionRate_1, ablateTemp
        integer(c_int32_t) :: loops, pulsesToDo, pulsetype, pad_
    end type tamc_heat_params

    integer(c_int), parameter :: TAMC_HEAT_TEMP = 0, TAMC_HEAT_RHOKAP = 1, TAMC_HEAT_WATER = 7, TAMC_HEAT_Q = 8, &
                                 TAMC_HEAT_TISSUE = 9, TAMC_HEAT_THRESTIME = 10, TAMC_HEAT_JMEAN = 11

    interface

        integer(c_int) function tamc_init(device, nxg, nyg, nzg, xmax, ymax, zmax, delta, handle) &
                bind(C, name="tamc_init")
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: device, nxg, nyg, nzg
            real(c_double), value :: xmax, ymax, zmax, delta
            type(c_ptr), intent(out) :: handle
        end function tamc_init

        integer(c_int) function tamc_finalize(handle) bind(C, name="tamc_finalize")
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function tamc_finalize

        integer(c_int) function tamc_set_source_co2(handle, spot_diameter_cm) bind(C, name="tamc_set_source_co2")
            import :: c_int, c_double, c_ptr
            type(c_ptr), value    :: handle
            real(c_double), value :: spot_diameter_cm
        end function tamc_set_source_co2

        !  Gaussian beam through rang (sourceph.f90:73-101) instead of the CO2 disk
        integer(c_int) function tamc_set_source_gaussian(handle, sigma_cm) bind(C, name="tamc_set_source_gaussian")
            import :: c_int, c_double, c_ptr
            type(c_ptr), value    :: handle
            real(c_double), value :: sigma_cm
        end function tamc_set_source_gaussian

        !  rhokap is passed as the whole allocatable rhokap(0:nxg+1,0:nyg+1,0:nzg+1): contiguous, so the
        !  compiler hands over the address of rhokap(0,0,0) without a temporary.
        integer(c_int) function tamc_set_optics(handle, rhokap, albedo, hgg, n1, n2, flags) &
                bind(C, name="tamc_set_optics")
            import :: c_int, c_double, c_ptr
            type(c_ptr), value        :: handle
            real(c_double), intent(in) :: rhokap(*)
            real(c_double), value     :: albedo, hgg, n1, n2
            integer(c_int), value     :: flags
        end function tamc_set_optics

        !  EXTENSION (no upstream counterpart): per-voxel albedo / hgg / refractive index, each an array shaped like rhokap
        !  or c_null_ptr for the scalar -- hence type(c_ptr) arguments (c_loc of a TARGET array)
        integer(c_int) function tamc_set_optics_grids(handle, albedo, hgg, n) bind(C, name="tamc_set_optics_grids")
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, albedo, hgg, n
        end function tamc_set_optics_grids

        integer(c_int) function tamc_run(handle, nphotons, seed, jmean_global, stats) bind(C, name="tamc_run")
            import :: c_int, c_int64_t, c_double, c_ptr
            type(c_ptr), value           :: handle
            integer(c_int64_t), value    :: nphotons, seed
            real(c_double), intent(out)  :: jmean_global(*)
            type(c_ptr), value           :: stats        ! c_loc(a tamc_stats) or c_null_ptr
        end function tamc_run

        !  tamc_set_optics + tamc_run in one call: what the time loop does once per iteration.  With rhokap and
        !  jmean_global page-locked (tamc_pin_host) the PCIe copies overlap the transport.
        integer(c_int) function tamc_run_optics(handle, rhokap, albedo, hgg, n1, n2, flags, nphotons, seed, &
                                                jmean_global, stats) bind(C, name="tamc_run_optics")
            import :: c_int, c_int64_t, c_double, c_ptr
            type(c_ptr), value           :: handle
            real(c_double), intent(in)   :: rhokap(*)
            real(c_double), value        :: albedo, hgg, n1, n2
            integer(c_int), value        :: flags
            integer(c_int64_t), value    :: nphotons, seed
            real(c_double), intent(out)  :: jmean_global(*)
            type(c_ptr), value           :: stats        ! c_loc(a tamc_stats) or c_null_ptr
        end function tamc_run_optics

        integer(c_int) function tamc_comm_unique_id(id128) bind(C, name="tamc_comm_unique_id")
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id128(128)
        end function tamc_comm_unique_id

        integer(c_int) function tamc_comm_init(handle, nranks, rank, id128) bind(C, name="tamc_comm_init")
            import :: c_int, c_char, c_ptr
            type(c_ptr), value     :: handle
            integer(c_int), value  :: nranks, rank
            character(kind=c_char), intent(in) :: id128(128)
        end function tamc_comm_init

        !  the array itself (assumed size, passed by reference): no c_loc, so iarray.f90's plain `real, allocatable`
        !  arrays need no TARGET attribute
        integer(c_int) function tamc_pin_host(a, bytes) bind(C, name="tamc_pin_host")
            import :: c_int, c_int64_t, c_double
            real(c_double), intent(in) :: a(*)
            integer(c_int64_t), value  :: bytes
        end function tamc_pin_host

        !  tuning / behaviour knobs by name (include/tamc.h); the name is a NUL-terminated string: 'root_io'//c_null_char
        integer(c_int) function tamc_set_option(handle, name, value) bind(C, name="tamc_set_option")
            import :: c_int, c_int64_t, c_char, c_ptr
            type(c_ptr), value                 :: handle
            character(kind=c_char), intent(in) :: name(*)
            integer(c_int64_t), value          :: value
        end function tamc_set_option

        integer(c_int) function tamc_device_count() bind(C, name="tamc_device_count")
            import :: c_int
        end function tamc_device_count

        !  ---- optional: the heat / ablation step on the device (replaces heat_sim_3d + arrhenius +
        !  setupThermalCoeff, mcpolar.f90:174-182) and the whole time loop resident on the GPU (:148-186)
        integer(c_int) function tamc_heat_init(handle, params, delt) bind(C, name="tamc_heat_init")
            import :: c_int, c_double, c_ptr, tamc_heat_params
            type(c_ptr), value          :: handle
            type(tamc_heat_params), intent(in) :: params
            real(c_double), intent(out) :: delt
        end function tamc_heat_init

        integer(c_int) function tamc_heat_step(handle, nphotons_times_numproc) bind(C, name="tamc_heat_step")
            import :: c_int, c_int64_t, c_ptr
            type(c_ptr), value        :: handle
            integer(c_int64_t), value :: nphotons_times_numproc
        end function tamc_heat_step

        integer(c_int) function tamc_coupled_loop(handle, nphotons, seed, max_iterations, iterations_done, packets_done) &
                bind(C, name="tamc_coupled_loop")
            import :: c_int, c_int64_t, c_ptr
            type(c_ptr), value              :: handle
            integer(c_int64_t), value       :: nphotons, seed, max_iterations
            integer(c_int64_t), intent(out) :: iterations_done, packets_done
        end function tamc_coupled_loop

        !  which = TAMC_HEAT_TEMP ... (include/tamc.h); upload = 0 copies device -> host
        integer(c_int) function tamc_heat_array(handle, which, host, upload) bind(C, name="tamc_heat_array")
            import :: c_int, c_double, c_ptr
            type(c_ptr), value    :: handle
            integer(c_int), value :: which, upload
            real(c_double)        :: host(*)
        end function tamc_heat_array

        function tamc_last_error() bind(C, name="tamc_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function tamc_last_error

    end interface

    contains

        subroutine tamc_check(ierr, where)
        !  reference convention on failure is print + error stop (inttau2.f90:171-172, 3dFD.f90:49)
            integer(c_int),   intent(in) :: ierr
            character(len=*), intent(in) :: where

            character(kind=c_char), pointer :: cmsg(:)
            integer :: n

            if(ierr == TAMC_OK)return
            call c_f_pointer(tamc_last_error(), cmsg, [512])
            n = 1
            do while(n < 512 .and. cmsg(n) /= c_null_char)
                n = n + 1
            end do
            print*, 'libtamc error ', ierr, ' in ', where, ': ', cmsg(1:n-1)
            error stop 1
        end subroutine tamc_check

end module tamc_mod
