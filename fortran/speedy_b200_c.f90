!> iso_c_binding interfaces of libspeedy_b200.so — EVERY entry point of include/speedy_b200.h — and thin replacements of the
!  reference's hot-path modules (legendre, fourier, spectral, tendencies, physics, time_stepping) that keep their public names
!  and argument lists.  SOURCE ONLY: this image has no Fortran compiler, so this file is not built here (DESIGN.md section 5).
!  What CAN be checked without one is: the C side of the derived type (tests/helpers/abi_layout.c asserts sizeof / offsetof of
!  speedy_cfg against the sequence-associated layout a bind(C) type has) and the call sequence below driven by reference from
!  compiled C exactly as these wrappers pass their arguments (tests/test_abi_c.py).
module speedy_b200_c
    use, intrinsic :: iso_c_binding
    implicit none

    type, bind(C) :: speedy_cfg
        integer(c_int) :: trunc, kx, ntr, nmembers, device, sppt_on
        integer(c_long_long) :: seed
        integer(c_int) :: member_offset, nsteps, precision
    end type

    !> namelist.nml as the reference reads it (params.f90:46-70, date.f90:54-71); datetimes = (year, month, day, hour, minute)
    type, bind(C) :: speedy_namelist
        integer(c_int) :: nsteps_out, nstdia
        integer(c_int) :: start_datetime(5), end_datetime(5)
    end type

    interface
        integer(c_int) function speedy_create(cfg, ctx) bind(C, name="speedy_create")
            import; type(speedy_cfg), intent(in) :: cfg; type(c_ptr), intent(out) :: ctx
        end function
        integer(c_int) function speedy_destroy(ctx) bind(C, name="speedy_destroy")
            import; type(c_ptr), value :: ctx
        end function
        function speedy_last_error() bind(C, name="speedy_last_error") result(msg)
            import; type(c_ptr) :: msg
        end function
        integer(c_int) function speedy_model_init(ctx, bc_path, y, m, d, h, mi) bind(C, name="speedy_model_init")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: bc_path(*)
            integer(c_int), value :: y, m, d, h, mi
        end function
        ! spectral.f90:98 / :112
        integer(c_int) function speedy_spec_to_grid(ctx, spec, nbatch, kcos, grid) bind(C, name="speedy_spec_to_grid")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: spec(*)
            integer(c_int), value :: nbatch; integer(c_int), intent(in) :: kcos(*); real(c_double), intent(out) :: grid(*)
        end function
        integer(c_int) function speedy_grid_to_spec(ctx, grid, nbatch, spec) bind(C, name="speedy_grid_to_spec")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: grid(*)
            integer(c_int), value :: nbatch; complex(c_double_complex), intent(out) :: spec(*)
        end function
        ! legendre.f90:74 / :114, fourier.f90:23 / :56
        integer(c_int) function speedy_legendre_inv(ctx, a, nbatch, b) bind(C, name="speedy_legendre_inv")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: a(*); integer(c_int), value :: nbatch; real(c_double), intent(out) :: b(*)
        end function
        integer(c_int) function speedy_legendre_dir(ctx, a, nbatch, b) bind(C, name="speedy_legendre_dir")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: a(*); integer(c_int), value :: nbatch; real(c_double), intent(out) :: b(*)
        end function
        integer(c_int) function speedy_fourier_inv(ctx, a, nbatch, kcos, b) bind(C, name="speedy_fourier_inv")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: a(*); integer(c_int), value :: nbatch
            integer(c_int), intent(in) :: kcos(*); real(c_double), intent(out) :: b(*)
        end function
        integer(c_int) function speedy_fourier_dir(ctx, a, nbatch, b) bind(C, name="speedy_fourier_dir")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: a(*); integer(c_int), value :: nbatch; real(c_double), intent(out) :: b(*)
        end function
        ! spectral.f90:124-233
        integer(c_int) function speedy_uvspec(ctx, vorm, divm, nbatch, ucosm, vcosm) bind(C, name="speedy_uvspec")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: vorm(*), divm(*)
            integer(c_int), value :: nbatch; complex(c_double_complex), intent(out) :: ucosm(*), vcosm(*)
        end function
        integer(c_int) function speedy_vdspec(ctx, ug, vg, nbatch, kcos, vorm, divm) bind(C, name="speedy_vdspec")
            import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: ug(*), vg(*)
            integer(c_int), value :: nbatch, kcos; complex(c_double_complex), intent(out) :: vorm(*), divm(*)
        end function
        ! module state by name (prognostics.f90:16-24, mod_radcon, auxiliaries, land/sea models)
        integer(c_int) function speedy_set_field(ctx, name, host, n) bind(C, name="speedy_set_field")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(in) :: host(*); integer(c_size_t), value :: n
        end function
        integer(c_int) function speedy_get_field(ctx, name, host, n) bind(C, name="speedy_get_field")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(out) :: host(*); integer(c_size_t), value :: n
        end function
        ! time_stepping.f90:12 / :35, implicit.f90:36
        integer(c_int) function speedy_initialize_implicit(ctx, dt) bind(C, name="speedy_initialize_implicit")
            import; type(c_ptr), value :: ctx; real(c_double), value :: dt
        end function
        integer(c_int) function speedy_first_step(ctx) bind(C, name="speedy_first_step")
            import; type(c_ptr), value :: ctx
        end function
        integer(c_int) function speedy_step(ctx, j1, j2, dt, compute_shortwave) bind(C, name="speedy_step")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: j1, j2, compute_shortwave; real(c_double), value :: dt
        end function
        ! speedy.f90:27-54 repeated; state resident on the device / on the host
        integer(c_int) function speedy_run_steps(ctx, nsteps) bind(C, name="speedy_run_steps")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: nsteps
        end function
        integer(c_int) function speedy_run_steps_host(ctx, state, n, nsteps, out) bind(C, name="speedy_run_steps_host")
            import; type(c_ptr), value :: ctx; real(c_double), intent(inout) :: state(*)
            integer(c_size_t), value :: n; integer(c_int), value :: nsteps; type(c_ptr), value :: out
        end function
        integer(c_int) function speedy_check_diagnostics(ctx, time_level, diag) bind(C, name="speedy_check_diagnostics")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: time_level; real(c_double), intent(out) :: diag(*)
        end function
        integer(c_int) function speedy_couple_sea_land(ctx, day) bind(C, name="speedy_couple_sea_land")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: day
        end function
        integer(c_int) function speedy_set_forcing(ctx, imode) bind(C, name="speedy_set_forcing")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: imode
        end function
        integer(c_int) function speedy_output_fields(ctx, member, u, v, t, q, phi, ps) bind(C, name="speedy_output_fields")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: member
            real(c_float), intent(out) :: u(*), v(*), t(*), q(*), phi(*), ps(*)
        end function
        ! output() file (input_output.f90:95-217) written by the library, and restart files; dir / path are c_null_char-terminated
        integer(c_int) function speedy_write_output(ctx, member, dir, path_out, path_cap) bind(C, name="speedy_write_output")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: member
            character(kind=c_char), intent(in) :: dir(*); type(c_ptr), value :: path_out; integer(c_size_t), value :: path_cap
        end function
        ! asynchronous form: conversions enqueued, files written by host threads; ymdhm / timestep = date and step of the enqueued state
        integer(c_int) function speedy_write_output_async(ctx, member, dir, ymdhm, timestep) bind(C, name="speedy_write_output_async")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: member; character(kind=c_char), intent(in) :: dir(*)
            integer(c_int), intent(in) :: ymdhm(5); integer(c_long_long), value :: timestep
        end function
        integer(c_int) function speedy_output_drain(ctx) bind(C, name="speedy_output_drain")
            import; type(c_ptr), value :: ctx
        end function
        integer(c_int) function speedy_save_restart(ctx, path) bind(C, name="speedy_save_restart")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: path(*)
        end function
        integer(c_int) function speedy_load_restart(ctx, path) bind(C, name="speedy_load_restart")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: path(*)
        end function
        ! ensembles: sppt.f90:45-99 noise source, on-device moments of the 41 output levels (device pointers)
        integer(c_int) function speedy_set_sppt_draw(ctx, on) bind(C, name="speedy_set_sppt_draw")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: on
        end function
        integer(c_int) function speedy_ensemble_sums_dev(ctx, d_sum, d_sumsq) bind(C, name="speedy_ensemble_sums_dev")
            import; type(c_ptr), value :: ctx, d_sum, d_sumsq
        end function
        integer(c_int) function speedy_model_date(ctx, ymdhm, model_step) bind(C, name="speedy_model_date")
            import; type(c_ptr), value :: ctx; integer(c_int), intent(out) :: ymdhm(5); integer(c_long_long), intent(out) :: model_step
        end function
        ! ---- life cycle / tables -------------------------------------------------------------------------------------
        integer(c_int) function speedy_synchronize(ctx) bind(C, name="speedy_synchronize")
            import; type(c_ptr), value :: ctx
        end function
        integer(c_int) function speedy_dims(ctx, dims) bind(C, name="speedy_dims")
            import; type(c_ptr), value :: ctx; integer(c_int), intent(out) :: dims(8)
        end function
        integer(c_int) function speedy_get_table(ctx, name, out, n) bind(C, name="speedy_get_table")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(out) :: out(*); integer(c_size_t), value :: n
        end function
        ! a Fortran host may hand over the tables its own initialize_* computed (table arithmetic shared bit for bit)
        integer(c_int) function speedy_set_table(ctx, name, tab, n) bind(C, name="speedy_set_table")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(in) :: tab(*); integer(c_size_t), value :: n
        end function
        integer(c_int) function speedy_host_table(trunc, name, out, n) bind(C, name="speedy_host_table")
            import; integer(c_int), value :: trunc; character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(out) :: out(*); integer(c_size_t), value :: n
        end function
        integer(c_long_long) function speedy_host_table_len(trunc, name) bind(C, name="speedy_host_table_len")
            import; integer(c_int), value :: trunc; character(kind=c_char), intent(in) :: name(*)
        end function
        function speedy_field_names() bind(C, name="speedy_field_names") result(names)
            import; type(c_ptr) :: names
        end function
        ! ---- device-pointer transforms (enqueue only; kcos stays a host array) ------------------------------------------
        integer(c_int) function speedy_spec_to_grid_dev(ctx, d_spec, nbatch, kcos, d_grid) bind(C, name="speedy_spec_to_grid_dev")
            import; type(c_ptr), value :: ctx, d_spec, d_grid; integer(c_int), value :: nbatch; integer(c_int), intent(in) :: kcos(*)
        end function
        integer(c_int) function speedy_grid_to_spec_dev(ctx, d_grid, nbatch, d_spec) bind(C, name="speedy_grid_to_spec_dev")
            import; type(c_ptr), value :: ctx, d_grid, d_spec; integer(c_int), value :: nbatch
        end function
        ! ---- spectral operators spectral.f90:84-96,124-171,229 -----------------------------------------------------------
        integer(c_int) function speedy_laplacian(ctx, a, nbatch, b) bind(C, name="speedy_laplacian")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: a(*); integer(c_int), value :: nbatch
            complex(c_double_complex), intent(out) :: b(*)
        end function
        integer(c_int) function speedy_inverse_laplacian(ctx, a, nbatch, b) bind(C, name="speedy_inverse_laplacian")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: a(*); integer(c_int), value :: nbatch
            complex(c_double_complex), intent(out) :: b(*)
        end function
        integer(c_int) function speedy_grad(ctx, psi, nbatch, psdx, psdy) bind(C, name="speedy_grad")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: psi(*); integer(c_int), value :: nbatch
            complex(c_double_complex), intent(out) :: psdx(*), psdy(*)
        end function
        integer(c_int) function speedy_vds(ctx, ucosm, vcosm, nbatch, vorm, divm) bind(C, name="speedy_vds")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: ucosm(*), vcosm(*)
            integer(c_int), value :: nbatch; complex(c_double_complex), intent(out) :: vorm(*), divm(*)
        end function
        integer(c_int) function speedy_trunct(ctx, vor, nbatch) bind(C, name="speedy_trunct")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(inout) :: vor(*); integer(c_int), value :: nbatch
        end function
        ! ---- module state, tendencies, physics --------------------------------------------------------------------------
        integer(c_int) function speedy_get_ifield(ctx, name, host, n) bind(C, name="speedy_get_ifield")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*)
            integer(c_int), intent(out) :: host(*); integer(c_size_t), value :: n
        end function
        ! implicit.f90:168-217 and horizontal_diffusion.f90:86-105 as stand-alone operators on host arrays (in place)
        integer(c_int) function speedy_implicit_terms(ctx, divdt, tdt, psdt) bind(C, name="speedy_implicit_terms")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(inout) :: divdt(*), tdt(*), psdt(*)
        end function
        integer(c_int) function speedy_do_horizontal_diffusion(ctx, field, fdt, dmp, dmp1, nlev) bind(C, name="speedy_do_horizontal_diffusion")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: field(*); complex(c_double_complex), intent(inout) :: fdt(*)
            real(c_double), intent(in) :: dmp(*), dmp1(*); integer(c_int), value :: nlev
        end function
        integer(c_int) function speedy_get_geopotential(ctx, j) bind(C, name="speedy_get_geopotential")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: j
        end function
        ! tendencies.f90:11-37 on the resident state; results in the fields "vordt","divdt","tdt","psdt","trdt"
        integer(c_int) function speedy_get_tendencies(ctx, j2, compute_shortwave) bind(C, name="speedy_get_tendencies")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: j2, compute_shortwave
        end function
        ! physics.f90:43-223 on host arrays
        integer(c_int) function speedy_get_physical_tendencies(ctx, vor, div, t, q, phi, psl, utend, vtend, ttend, qtend, compute_shortwave) &
                bind(C, name="speedy_get_physical_tendencies")
            import; type(c_ptr), value :: ctx; complex(c_double_complex), intent(in) :: vor(*), div(*), t(*), q(*), phi(*), psl(*)
            real(c_double), intent(inout) :: utend(*), vtend(*), ttend(*), qtend(*); integer(c_int), value :: compute_shortwave
        end function
        ! step(j1,j2,dt) with the prognostic arrays left on the host: state = [vor, div, t, tr, ps] (prognostics.f90:16-20) back to back
        integer(c_int) function speedy_step_host(ctx, state, n, j1, j2, dt, compute_shortwave) bind(C, name="speedy_step_host")
            import; type(c_ptr), value :: ctx; real(c_double), intent(inout) :: state(*); integer(c_size_t), value :: n
            integer(c_int), value :: j1, j2, compute_shortwave; real(c_double), value :: dt
        end function
        integer(c_size_t) function speedy_state_len(ctx) bind(C, name="speedy_state_len")
            import; type(c_ptr), value :: ctx
        end function
        integer(c_size_t) function speedy_output_len(ctx) bind(C, name="speedy_output_len")
            import; type(c_ptr), value :: ctx
        end function
        ! the main loop in two halves: enqueue only / drain + range guard
        integer(c_int) function speedy_enqueue_steps(ctx, nsteps) bind(C, name="speedy_enqueue_steps")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: nsteps
        end function
        integer(c_int) function speedy_finish(ctx) bind(C, name="speedy_finish")
            import; type(c_ptr), value :: ctx
        end function
        ! host-only writer of one output() file (input_output.f90:95-217)
        integer(c_int) function speedy_write_output_file(path, trunc, nsteps, start_ymdhm, timestep, u, v, t, q, phi, ps) bind(C, name="speedy_write_output_file")
            import; character(kind=c_char), intent(in) :: path(*); integer(c_int), value :: trunc, nsteps, timestep
            integer(c_int), intent(in) :: start_ymdhm(5); real(c_float), intent(in) :: u(*), v(*), t(*), q(*), phi(*), ps(*)
        end function
        ! ---- measurement / debugging aids ---------------------------------------------------------------------------------
        integer(c_long_long) function speedy_launch_count(ctx) bind(C, name="speedy_launch_count")
            import; type(c_ptr), value :: ctx
        end function
        function speedy_stream(ctx) bind(C, name="speedy_stream") result(stream)
            import; type(c_ptr), value :: ctx; type(c_ptr) :: stream
        end function
        integer(c_int) function speedy_set_graphs(ctx, on) bind(C, name="speedy_set_graphs")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: on
        end function
        integer(c_int) function speedy_set_option(ctx, name, val) bind(C, name="speedy_set_option")
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: name(*); integer(c_int), value :: val
        end function
        integer(c_int) function speedy_time_kernels(ctx, nsteps, flush_l2, ms) bind(C, name="speedy_time_kernels")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: nsteps, flush_l2; real(c_double), intent(out) :: ms(*)
        end function
        function speedy_kernel_names() bind(C, name="speedy_kernel_names") result(names)
            import; type(c_ptr) :: names
        end function
        integer(c_int) function speedy_trace(ctx, on) bind(C, name="speedy_trace")
            import; type(c_ptr), value :: ctx; integer(c_int), value :: on
        end function
        integer(c_int) function speedy_trace_read(ctx, out9) bind(C, name="speedy_trace_read")
            import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: out9(9)
        end function
        integer(c_int) function speedy_host_calendar(ymdhm, nsteps, tmonth, tyear, imont1) bind(C, name="speedy_host_calendar")
            import; integer(c_int), intent(inout) :: ymdhm(5); integer(c_int), value :: nsteps
            real(c_double), intent(out) :: tmonth, tyear; integer(c_int), intent(out) :: imont1
        end function
        ! ---- program speedy (speedy.f90:1-54) on the library: namelist.nml, trip count, main loop -------------------------
        integer(c_int) function speedy_namelist_defaults(nml) bind(C, name="speedy_namelist_defaults")
            import; type(speedy_namelist), intent(out) :: nml
        end function
        integer(c_int) function speedy_read_namelist(path, nml) bind(C, name="speedy_read_namelist")
            import; character(kind=c_char), intent(in) :: path(*); type(speedy_namelist), intent(out) :: nml
        end function
        integer(c_long_long) function speedy_steps_between(start_ymdhm, end_ymdhm, nsteps) bind(C, name="speedy_steps_between")
            import; integer(c_int), intent(in) :: start_ymdhm(5), end_ymdhm(5); integer(c_int), value :: nsteps
        end function
        integer(c_int) function speedy_run_info(ctx, info) bind(C, name="speedy_run_info")
            import; type(c_ptr), value :: ctx; integer(c_int), intent(out) :: info(4)
        end function
        integer(c_long_long) function speedy_host_boundary(bc_path, trunc, name, out, n) bind(C, name="speedy_host_boundary")
            import; character(kind=c_char), intent(in) :: bc_path(*), name(*); integer(c_int), value :: trunc
            real(c_double), intent(out) :: out(*); integer(c_size_t), value :: n
        end function
        integer(c_int) function speedy_range_failure(ctx, step, diag) bind(C, name="speedy_range_failure")
            import; type(c_ptr), value :: ctx; integer(c_long_long), intent(out) :: step; real(c_double), intent(out) :: diag(24)
        end function
        integer(c_int) function speedy_main_loop(ctx, nml, out_dir, member, verbose, steps_done) bind(C, name="speedy_main_loop")
            import; type(c_ptr), value :: ctx; type(speedy_namelist), intent(in) :: nml; character(kind=c_char), intent(in) :: out_dir(*)
            integer(c_int), value :: member, verbose; integer(c_long_long), intent(out) :: steps_done
        end function
    end interface

    type(c_ptr), save :: b200_ctx = c_null_ptr   !! one context per process, like the reference's module state

contains
    !> non-zero return -> the reference's `stop` convention
    subroutine b200_check(rc, what)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: what
        if (rc > 0) stop 'Model variables out of accepted range'
        if (rc < 0) then
            print *, 'speedy_b200: ', what, ' failed'
            stop
        end if
    end subroutine
end module

!> Drop-in for module `legendre` (legendre.f90:11-12): legendre_inv / legendre_dir on real(2*mx, .) arrays.
module legendre_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    function legendre_inv(input) result(output)
        real(p), intent(in) :: input(2*mx,nx)
        real(p) :: output(2*mx,il)
        call b200_check(speedy_legendre_inv(b200_ctx, input, 1_c_int, output), 'legendre_inv')
    end function
    function legendre_dir(input) result(output)
        real(p), intent(in) :: input(2*mx,il)
        real(p) :: output(2*mx,nx)
        call b200_check(speedy_legendre_dir(b200_ctx, input, 1_c_int, output), 'legendre_dir')
    end function
end module

!> Drop-in for module `fourier` (fourier.f90:11).
module fourier_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    function fourier_inv(input, kcos) result(output)
        real(p), intent(in) :: input(2*mx,il)
        integer, intent(in) :: kcos
        real(p) :: output(ix,il)
        integer(c_int) :: k(1)
        k(1) = kcos
        call b200_check(speedy_fourier_inv(b200_ctx, input, 1_c_int, k, output), 'fourier_inv')
    end function
    function fourier_dir(input) result(output)
        real(p), intent(in) :: input(ix,il)
        real(p) :: output(2*mx,il)
        call b200_check(speedy_fourier_dir(b200_ctx, input, 1_c_int, output), 'fourier_dir')
    end function
end module

!> Drop-in for module `spectral` (spectral.f90:8-11): same public names and signatures, transforms and operators.
module spectral_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    function spec_to_grid(vorm, kcos) result(vorg)
        complex(p), intent(in) :: vorm(mx,nx)
        integer, intent(in) :: kcos
        real(p) :: vorg(ix,il)
        integer(c_int) :: k(1)
        k(1) = kcos
        call b200_check(speedy_spec_to_grid(b200_ctx, vorm, 1_c_int, k, vorg), 'spec_to_grid')
    end function
    function grid_to_spec(vorg) result(vorm)
        real(p), intent(in) :: vorg(ix,il)
        complex(p) :: vorm(mx,nx)
        call b200_check(speedy_grid_to_spec(b200_ctx, vorg, 1_c_int, vorm), 'grid_to_spec')
    end function
    function laplacian(input) result(output)
        complex(p), intent(in) :: input(mx,nx)
        complex(p) :: output(mx,nx)
        call b200_check(speedy_laplacian(b200_ctx, input, 1_c_int, output), 'laplacian')
    end function
    function inverse_laplacian(input) result(output)
        complex(p), intent(in) :: input(mx,nx)
        complex(p) :: output(mx,nx)
        call b200_check(speedy_inverse_laplacian(b200_ctx, input, 1_c_int, output), 'inverse_laplacian')
    end function
    subroutine grad(psi, psdx, psdy)
        complex(p), intent(in) :: psi(mx,nx)
        complex(p), intent(inout) :: psdx(mx,nx), psdy(mx,nx)
        call b200_check(speedy_grad(b200_ctx, psi, 1_c_int, psdx, psdy), 'grad')
    end subroutine
    subroutine vds(ucosm, vcosm, vorm, divm)
        complex(p), intent(in) :: ucosm(mx,nx), vcosm(mx,nx)
        complex(p), intent(inout) :: vorm(mx,nx), divm(mx,nx)
        call b200_check(speedy_vds(b200_ctx, ucosm, vcosm, 1_c_int, vorm, divm), 'vds')
    end subroutine
    subroutine uvspec(vorm, divm, ucosm, vcosm)
        complex(p), intent(in) :: vorm(mx,nx), divm(mx,nx)
        complex(p), intent(inout) :: ucosm(mx,nx), vcosm(mx,nx)
        call b200_check(speedy_uvspec(b200_ctx, vorm, divm, 1_c_int, ucosm, vcosm), 'uvspec')
    end subroutine
    subroutine vdspec(ug, vg, vorm, divm, kcos)
        real(p), intent(in) :: ug(ix,il), vg(ix,il)
        complex(p), intent(out) :: vorm(mx,nx), divm(mx,nx)
        integer, intent(in) :: kcos
        call b200_check(speedy_vdspec(b200_ctx, ug, vg, 1_c_int, int(kcos, c_int), vorm, divm), 'vdspec')
    end subroutine
    subroutine trunct(vor)
        complex(p), intent(inout) :: vor(mx,nx)
        call b200_check(speedy_trunct(b200_ctx, vor, 1_c_int), 'trunct')
    end subroutine
end module

!> Pack / unpack the module arrays of `prognostics` (prognostics.f90:16-20) as the contiguous state vector the host-resident
!  entry points take: [vor, div, t, tr, ps], each with both time levels, Fortran order.
module prognostics_b200
    use types, only: p
    use params
    use speedy_b200_c
    use prognostics, only: vor, div, t, tr, ps
    implicit none
    integer, parameter :: n3 = 2*mx*nx*kx*2, n2 = 2*mx*nx*2     !! reals per 3-D field (two time levels) / per ps
contains
    subroutine pack_state(state)
        real(c_double), intent(out) :: state(4*n3*1 + n2 + n3*(ntr-1))
        integer :: o
        o = 0
        state(o+1:o+n3) = transfer(vor, state(1:n3)); o = o + n3
        state(o+1:o+n3) = transfer(div, state(1:n3)); o = o + n3
        state(o+1:o+n3) = transfer(t, state(1:n3));   o = o + n3
        state(o+1:o+n3*ntr) = transfer(tr, state(1:n3*ntr)); o = o + n3*ntr
        state(o+1:o+n2) = transfer(ps, state(1:n2))
    end subroutine
    subroutine unpack_state(state)
        real(c_double), intent(in) :: state(4*n3*1 + n2 + n3*(ntr-1))
        integer :: o
        o = 0
        vor = reshape(transfer(state(o+1:o+n3), vor), shape(vor)); o = o + n3
        div = reshape(transfer(state(o+1:o+n3), div), shape(div)); o = o + n3
        t   = reshape(transfer(state(o+1:o+n3), t), shape(t));     o = o + n3
        tr  = reshape(transfer(state(o+1:o+n3*ntr), tr), shape(tr)); o = o + n3*ntr
        ps  = reshape(transfer(state(o+1:o+n2), ps), shape(ps))
    end subroutine
end module

!> Drop-in for module `time_stepping` (time_stepping.f90:8).  The prognostic module arrays stay authoritative on the HOST:
!  step() uploads them, runs step(j1,j2,dt) on the device and downloads them (speedy_step_host) — a literal replacement of
!  `call step`.  The production path is speedy_run_steps / speedy_run_steps_host, which amortise the copies over whole days.
module time_stepping_b200
    use types, only: p
    use params
    use speedy_b200_c
    use prognostics_b200
    implicit none
contains
    subroutine first_step
        real(c_double) :: state(4*n3 + n2 + n3*(ntr-1))
        ! first_step works on the resident state: push the host arrays, run, pull them back
        call pack_state(state)
        call b200_check(speedy_set_field(b200_ctx, 'vor'//c_null_char, state(1:n3), int(n3, c_size_t)), 'set vor')
        call b200_check(speedy_set_field(b200_ctx, 'div'//c_null_char, state(n3+1:2*n3), int(n3, c_size_t)), 'set div')
        call b200_check(speedy_set_field(b200_ctx, 't'//c_null_char, state(2*n3+1:3*n3), int(n3, c_size_t)), 'set t')
        call b200_check(speedy_set_field(b200_ctx, 'tr'//c_null_char, state(3*n3+1:3*n3+n3*ntr), int(n3*ntr, c_size_t)), 'set tr')
        call b200_check(speedy_set_field(b200_ctx, 'ps'//c_null_char, state(3*n3+n3*ntr+1:), int(n2, c_size_t)), 'set ps')
        call b200_check(speedy_first_step(b200_ctx), 'first_step')
        call b200_check(speedy_get_field(b200_ctx, 'vor'//c_null_char, state(1:n3), int(n3, c_size_t)), 'get vor')
        call b200_check(speedy_get_field(b200_ctx, 'div'//c_null_char, state(n3+1:2*n3), int(n3, c_size_t)), 'get div')
        call b200_check(speedy_get_field(b200_ctx, 't'//c_null_char, state(2*n3+1:3*n3), int(n3, c_size_t)), 'get t')
        call b200_check(speedy_get_field(b200_ctx, 'tr'//c_null_char, state(3*n3+1:3*n3+n3*ntr), int(n3*ntr, c_size_t)), 'get tr')
        call b200_check(speedy_get_field(b200_ctx, 'ps'//c_null_char, state(3*n3+n3*ntr+1:), int(n2, c_size_t)), 'get ps')
        call unpack_state(state)
    end subroutine
    subroutine step(j1, j2, dt)
        use shortwave_radiation, only: compute_shortwave
        integer, intent(in) :: j1, j2
        real(p), intent(in) :: dt
        real(c_double) :: state(4*n3 + n2 + n3*(ntr-1))
        call pack_state(state)
        call b200_check(speedy_step_host(b200_ctx, state, int(size(state), c_size_t), int(j1, c_int), int(j2, c_int), real(dt, c_double), &
                                         merge(1_c_int, 0_c_int, compute_shortwave)), 'step')
        call unpack_state(state)
    end subroutine
end module

!> Drop-in for module `tendencies` (tendencies.f90:8): get_tendencies on the host arrays of `prognostics`.
module tendencies_b200
    use types, only: p
    use params
    use speedy_b200_c
    use prognostics_b200
    implicit none
contains
    subroutine get_tendencies(vordt, divdt, tdt, psdt, trdt, j2)
        use shortwave_radiation, only: compute_shortwave
        complex(p), intent(inout) :: vordt(mx,nx,kx), divdt(mx,nx,kx), tdt(mx,nx,kx), psdt(mx,nx), trdt(mx,nx,kx,ntr)
        integer, intent(in) :: j2
        real(c_double) :: state(4*n3 + n2 + n3*(ntr-1)), buf(2*mx*nx*kx)
        call pack_state(state)
        call b200_check(speedy_set_field(b200_ctx, 'vor'//c_null_char, state(1:n3), int(n3, c_size_t)), 'set vor')
        call b200_check(speedy_set_field(b200_ctx, 'div'//c_null_char, state(n3+1:2*n3), int(n3, c_size_t)), 'set div')
        call b200_check(speedy_set_field(b200_ctx, 't'//c_null_char, state(2*n3+1:3*n3), int(n3, c_size_t)), 'set t')
        call b200_check(speedy_set_field(b200_ctx, 'tr'//c_null_char, state(3*n3+1:3*n3+n3*ntr), int(n3*ntr, c_size_t)), 'set tr')
        call b200_check(speedy_set_field(b200_ctx, 'ps'//c_null_char, state(3*n3+n3*ntr+1:), int(n2, c_size_t)), 'set ps')
        call b200_check(speedy_get_tendencies(b200_ctx, int(j2, c_int), merge(1_c_int, 0_c_int, compute_shortwave)), 'get_tendencies')
        call b200_check(speedy_get_field(b200_ctx, 'vordt'//c_null_char, buf, int(size(buf), c_size_t)), 'get vordt')
        vordt = reshape(transfer(buf, vordt), shape(vordt))
        call b200_check(speedy_get_field(b200_ctx, 'divdt'//c_null_char, buf, int(size(buf), c_size_t)), 'get divdt')
        divdt = reshape(transfer(buf, divdt), shape(divdt))
        call b200_check(speedy_get_field(b200_ctx, 'tdt'//c_null_char, buf, int(size(buf), c_size_t)), 'get tdt')
        tdt = reshape(transfer(buf, tdt), shape(tdt))
        call b200_check(speedy_get_field(b200_ctx, 'trdt'//c_null_char, buf, int(size(buf), c_size_t)), 'get trdt')
        trdt(:,:,:,1) = reshape(transfer(buf, trdt(:,:,:,1)), shape(trdt(:,:,:,1)))
        call b200_check(speedy_get_field(b200_ctx, 'psdt'//c_null_char, buf(1:2*mx*nx), int(2*mx*nx, c_size_t)), 'get psdt')
        psdt = reshape(transfer(buf(1:2*mx*nx), psdt), shape(psdt))
    end subroutine
end module

!> Drop-in for the procedures of module `implicit` (implicit.f90:11-12): initialize_implicit(dt), implicit_terms(divdt, tdt, psdt).
module implicit_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    subroutine initialize_implicit(dt)
        real(p), intent(in) :: dt
        call b200_check(speedy_initialize_implicit(b200_ctx, real(dt, c_double)), 'initialize_implicit')
    end subroutine
    subroutine implicit_terms(divdt, tdt, psdt)
        complex(p), intent(inout) :: divdt(mx,nx,kx), tdt(mx,nx,kx), psdt(mx,nx)
        call b200_check(speedy_implicit_terms(b200_ctx, divdt, tdt, psdt), 'implicit_terms')
    end subroutine
end module

!> Drop-in for the generic `do_horizontal_diffusion` of module `horizontal_diffusion` (horizontal_diffusion.f90:13-17,86-105).
module horizontal_diffusion_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
    interface do_horizontal_diffusion
        module procedure do_horizontal_diffusion_2d
        module procedure do_horizontal_diffusion_3d
    end interface
contains
    function do_horizontal_diffusion_2d(field, fdt_in, dmp, dmp1) result(fdt_out)
        complex(p), intent(in) :: field(mx,nx), fdt_in(mx,nx)
        real(p), intent(in) :: dmp(mx,nx), dmp1(mx,nx)
        complex(p) :: fdt_out(mx,nx)
        fdt_out = fdt_in
        call b200_check(speedy_do_horizontal_diffusion(b200_ctx, field, fdt_out, dmp, dmp1, 1_c_int), 'do_horizontal_diffusion')
    end function
    function do_horizontal_diffusion_3d(field, fdt_in, dmp, dmp1) result(fdt_out)
        complex(p), intent(in) :: field(mx,nx,kx), fdt_in(mx,nx,kx)
        real(p), intent(in) :: dmp(mx,nx), dmp1(mx,nx)
        complex(p) :: fdt_out(mx,nx,kx)
        fdt_out = fdt_in
        call b200_check(speedy_do_horizontal_diffusion(b200_ctx, field, fdt_out, dmp, dmp1, int(kx, c_int)), 'do_horizontal_diffusion')
    end function
end module

!> Drop-in for module `physics` (physics.f90:8,43): get_physical_tendencies on host arrays.
module physics_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    subroutine get_physical_tendencies(vor, div, t, q, phi, psl, utend, vtend, ttend, qtend)
        use shortwave_radiation, only: compute_shortwave
        complex(p), intent(in) :: vor(mx,nx,kx), div(mx,nx,kx), t(mx,nx,kx), q(mx,nx,kx), phi(mx,nx,kx), psl(mx,nx)
        real(p), intent(inout) :: utend(ix,il,kx), vtend(ix,il,kx), ttend(ix,il,kx), qtend(ix,il,kx)
        call b200_check(speedy_get_physical_tendencies(b200_ctx, vor, div, t, q, phi, psl, utend, vtend, ttend, qtend, &
                                                       merge(1_c_int, 0_c_int, compute_shortwave)), 'get_physical_tendencies')
    end subroutine
end module

!> Drop-in for `program speedy` (speedy.f90:1-54) when nothing of the reference but its namelist.nml and boundary files is kept: the
!> five calls `speedy.f90_b200/bin/speedy_b200` makes from C++.  Build this unit with -cpp -DSPEEDY_B200_PROGRAM (it is the only `program`
!> of the file; the modules above are for a host that keeps the reference's own main program).
#ifdef SPEEDY_B200_PROGRAM
program speedy_b200_main
    use speedy_b200_c
    implicit none
    type(speedy_cfg) :: cfg
    type(speedy_namelist) :: nml
    integer(c_long_long) :: steps
    integer(c_int) :: rc

    call b200_check(speedy_read_namelist('namelist.nml'//c_null_char, nml), 'namelist')            ! params.f90:54-70, date.f90:54-71
    cfg = speedy_cfg(30_c_int, 8_c_int, 1_c_int, 1_c_int, 0_c_int, 0_c_int, 0_c_long_long, 0_c_int, 0_c_int, 0_c_int)
    call b200_check(speedy_create(cfg, b200_ctx), 'create')
    ! initialize (initialization.f90:12-82) from the boundary files of the run directory, start date of the namelist
    call b200_check(speedy_model_init(b200_ctx, '.'//c_null_char, nml%start_datetime(1), nml%start_datetime(2), nml%start_datetime(3), &
                                      nml%start_datetime(4), nml%start_datetime(5)), 'initialize')
    ! speedy.f90:24-54: files every nsteps_out steps into the working directory, diagnostics every nstdia steps
    rc = speedy_main_loop(b200_ctx, nml, '.'//c_null_char, 0_c_int, 1_c_int, steps)
    call b200_check(rc, 'main loop')                                                               ! rc = 1: stop 'Model variables out of accepted range'
    call b200_check(speedy_destroy(b200_ctx), 'destroy')
end program
#endif
