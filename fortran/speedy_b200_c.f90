!> iso_c_binding interfaces of libspeedy_b200.so (include/speedy_b200.h) and thin
!  replacements of the reference's hot-path procedures that keep their names and argument
!  lists.  SOURCE ONLY: this image has no Fortran compiler, so this file is not built or
!  tested here (DESIGN.md section 5); the same C ABI is exercised from C++/Python tests.
module speedy_b200_c
    use, intrinsic :: iso_c_binding
    implicit none

    type, bind(C) :: speedy_cfg
        integer(c_int) :: trunc, kx, ntr, nmembers, device, sppt_on
        integer(c_long_long) :: seed
        integer(c_int) :: member_offset, nsteps, precision
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

!> Drop-in for module `spectral` (spectral.f90:8-11): same public names and signatures.
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
end module

!> Drop-in for module `time_stepping` (time_stepping.f90:8): the prognostic module arrays stay in
!  `prognostics`; step() moves them through the device (speedy_run_steps_host does the same for
!  whole days and is what the main loop should call to amortise the copies).
module time_stepping_b200
    use types, only: p
    use params
    use speedy_b200_c
    implicit none
contains
    subroutine first_step
        call b200_check(speedy_first_step(b200_ctx), 'first_step')
    end subroutine
    subroutine step(j1, j2, dt)
        use shortwave_radiation, only: compute_shortwave
        integer, intent(in) :: j1, j2
        real(p), intent(in) :: dt
        call b200_check(speedy_step(b200_ctx, j1, j2, dt, merge(1_c_int, 0_c_int, compute_shortwave)), 'step')
    end subroutine
end module
