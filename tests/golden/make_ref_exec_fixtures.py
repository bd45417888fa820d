#!/usr/bin/env python
"""Generates tests/golden/ref_exec/*.npz: golden vectors produced by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCE
(src_compressible/{mhdinit,dealiasing,AEBmod,rktmod,mhdrhs,fftw,parallel,mhd,mhdrms}.f90, read where it lies under
/root/reference) through the statement-by-statement translator of oracle/fortran_exec.py, on one rank.

Run from the reference's text: grid_initialize, dealias_initialize, AEB_calc / evolve_radius / update_ksquare,
initial_calc_conserve_variable, transform_uu_real_to_fourier, vardt (+ rkt_init), evolve = 3 x { calc_flux
(+ calc_current_density_real), transform_flux_real_to_fourier, calc_rhs, rkt, dealias, transform_uu_fourier_to_real,
update_uu_prim_from_uu }, from_xyz_to_zxy / from_zxy_to_xyz and the single-rank branches of transpose_xy/yx/yz/zy,
calc_max_divB, calc_rms.  Supplied from outside: the three FFTW executions (numpy.fft on one line — the DFT FFTW
computes, unnormalised, c2r ignoring the imaginary parts of the DC and Nyquist bins) and mpi_allreduce on one rank.

Run in the build container (needs /root/reference):  python tests/golden/make_ref_exec_fixtures.py
tests/test_reference_source_pins.py then checks the oracle (CPU) and the library (emulator / GPU) against the stored
vectors without the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fortran_exec as fx  # noqa: E402

REFROOT = "/root/reference"
REF = REFROOT + "/src_compressible"

CASES = {
    # name: grid, namelist-level switches (3D compressible tree unless `tree` says otherwise)
    "hall_aeb_mask": dict(nx=32, ny=16, nz=16, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1,
                          if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
    "corot_filter_explicit": dict(nx=16, ny=32, nz=8, if_hall=True, if_aeb=True, if_corotating=True, dealias_option=2,
                                  if_resis=True, if_resis_exp=True, if_visc=True, if_visc_exp=True, if_conserve_background=True),
    # plain MHD: no Hall term, no expanding box, no dealiasing (the library then does the reference's 19 transforms)
    "plain_nodealias": dict(nx=16, ny=16, nz=16, if_hall=False, if_aeb=False, if_corotating=False, dealias_option=0,
                            if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
    # a line length with an odd factor (FFTW plans any length, fftw.f90:27-33): n/2+1 wave numbers, the 1/3 mask and the
    # library's composite transforms (fft_core.cuh) on the x axis
    "lines48_hall_aeb_corot_mask": dict(nx=48, ny=16, nz=16, if_hall=True, if_aeb=True, if_corotating=True, dealias_option=1,
                                        if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
}
# src_incompressible: pressure projection, J and grad u, update_rho_p
CASES_INCOMPRESSIBLE = {
    "incomp_hall_aeb_mask": dict(nx=16, ny=32, nz=16, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1,
                                 if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
    "incomp_corot_filter_explicit": dict(nx=32, ny=16, nz=8, if_hall=True, if_aeb=True, if_corotating=True, dealias_option=2,
                                         if_resis=True, if_resis_exp=True, if_visc=True, if_visc_exp=True, if_conserve_background=True),
    # no dealiasing: the spectrum is re-derived from the real fields at the start of every stage (mhd.f90:325), which matters here
    "incomp_plain_nodealias": dict(nx=16, ny=16, nz=16, if_hall=False, if_aeb=False, if_corotating=False, dealias_option=0,
                                   if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
    # 48-point lines on the y axis (3 * 16)
    "incomp_lines48_hall_aeb_mask": dict(nx=16, ny=48, nz=16, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1,
                                         if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False),
}


def build_namespace(c, incompressible=False):
    nx, ny, nz = c["nx"], c["ny"], c["nz"]
    nxh = nx // 2 + 1
    ns = fx.base_namespace()
    F = fx.FArray

    def real(*shape):      # C-ordered storage [.., z, y, x] seen by Fortran as (x, y, z, ..)
        return np.zeros(shape[::-1])

    def cplx(*shape):
        return np.zeros(shape[::-1], dtype=np.complex128)

    st = dict(uu=real(nx, ny, nz, 8), uu_prim=real(nx, ny, nz, 4), flux=real(nx, ny, nz, 18), expand_term=real(nx, ny, nz, 1),
              current_density=real(nx, ny, nz, 3), uu_fourier=cplx(nxh, ny, nz, 8), flux_fourier=cplx(nxh, ny, nz, 18),
              fnl=cplx(nxh, ny, nz, 8), fnl_rk=cplx(nxh, ny, nz, 8), current_density_fourier=cplx(nxh, ny, nz, 3),
              expand_term_fourier=cplx(nxh, ny, nz, 1), k_square=real(nxh, ny, nz), w_xyz=cplx(nxh, ny, nz),
              w_yxz=cplx(nxh, ny, nz), w_zxy=cplx(nxh, ny, nz), xgrid=real(nx), ygrid=real(ny), zgrid=real(nz),
              wave_number_x=real(nx), wave_number_y=real(ny), wave_number_z=real(nz), fx_aux=real(nx), fx_aux_ft=cplx(nxh),
              fy_aux=cplx(ny), fy_aux_ft=cplx(ny), fz_aux=cplx(nz), fz_aux_ft=cplx(nz), filtx=real(nxh), filty=real(ny),
              filtz=real(nz), cc1=real(3), dd1=real(3), time_step=real(3), uu_ave=real(8), uu_square_ave=real(8),
              uu_ave_sum=real(8), uu_square_ave_sum=real(8), uu_rms=real(8), b0_ave=real(8), rho_u2=real(3), rho_u2_sum=real(3))
    if incompressible:     # src_incompressible/mhdinit.f90:5-6,126-190: nvarPrim = nflux = nfluxPressure = 3
        st.update(uu_prim=real(nx, ny, nz, 3), flux=real(nx, ny, nz, 3), flux_pressure=real(nx, ny, nz, 3),
                  grad_velocity=real(nx, ny, nz, 9), divb_arr=real(nx, ny, nz, 1), divv_arr=real(nx, ny, nz, 1),
                  flux_fourier=cplx(nxh, ny, nz, 3), flux_pressure_fourier=cplx(nxh, ny, nz, 3),
                  grad_velocity_fourier=cplx(nxh, ny, nz, 9), divb_arr_fourier=cplx(nxh, ny, nz, 1), divv_arr_fourier=cplx(nxh, ny, nz, 1))
    for k, v in st.items():
        ns[k] = F(v.T)
    ns["_storage"] = st
    ns["_wrap_ints"] = lambda: ns.update({k: fx.FInt(v) for k, v in ns.items() if type(v) is int})   # Fortran INTEGER semantics
    one = lambda v: F(np.array([v]))  # noqa: E731
    # parallel_start on one rank (parallel.f90:100-143): offsets 0, sizes = the whole axis
    ns.update(xi_offset=one(0), xi_size=one(nxh), yi_offset=one(0), yi_size=one(ny), yj_offset=one(0), yj_size=one(ny),
              zj_offset=one(0), zj_size=one(nz), dims=F(np.array([1, 1])), myid_i=0, myid_j=0, ipe=0, npe=1, ierr=0)
    # namelist values (mhd.input of SURVEY 8(d)) and module defaults (mhdinit.f90:5-54, dealiasing.f90:9-10, AEBmod.f90:10-12)
    ns.update(nx=nx, ny=ny, nz=nz, nvar=8, pi=3.141592653589793, lx=24.0, ly=12.0, lz=6.0, dx=0.0, dy=0.0, dz=0.0,
              adiabatic_index=1.666667, resistivity=1e-3 if c["if_resis_exp"] else 1e-4, viscosity=1e-3 if c["if_visc_exp"] else 1e-4,
              ion_inertial_length=0.2, cfl=0.5, afx=0.495, afy=0.495, afz=0.495, dealias_circle_radius=1. / 3.,
              radius0=30.0, radius=30.0, ur0=1.167, ur=0.0, tau_exp=0.0, corotating_angle=0.3 if c["if_corotating"] else 0.0,
              cos_cor_ang=1.0, sin_cor_ang=0.0, time=0.0, dt=0.0, max_divb=0.0, size_grid=nx * ny * nz,
              mpi_realtype=None, mpi_sum=None, mpi_min=None, mpi_max=None, mpi_comm_world=None)
    if incompressible:
        ns.update(nvarprim=3, nflux=3, nfluxpressure=3, rho0=1.0, p0=1.0, t0=1.0, max_divv=0.0)
    for k in ("if_hall", "if_aeb", "if_corotating", "dealias_option", "if_resis", "if_resis_exp", "if_visc", "if_visc_exp",
              "if_conserve_background"):
        ns[k] = c[k]
    ns.update(plan_fft_x="r2c", plan_ifft_x="c2r", plan_fft_y=-1, plan_ifft_y=+1, plan_fft_z=-1, plan_ifft_z=+1)

    # ---- the only pieces not taken from the reference's text: FFTW's 1-D executions (fftw.f90:27-33 plans)
    def fftw_execute_dft_r2c(plan, a, out):
        out.a[...] = np.fft.rfft(a.a)

    def fftw_execute_dft_c2r(plan, a, out):
        h = a.a.copy()
        h[0] = h[0].real
        h[-1] = h[-1].real                       # FFTW's c2r ignores the imaginary parts of the DC and Nyquist bins
        out.a[...] = np.fft.irfft(h, n=out.a.size) * out.a.size

    def fftw_execute_dft(plan, a, out):
        out.a[...] = np.fft.fft(a.a) if plan < 0 else np.fft.ifft(a.a) * a.a.size

    ns.update(fftw_execute_dft_r2c=fftw_execute_dft_r2c, fftw_execute_dft_c2r=fftw_execute_dft_c2r, fftw_execute_dft=fftw_execute_dft)
    ns["_wrap_ints"]()
    return ns


def load_reference(ns):
    src = {}
    src.update(fx.load(ns, f"{REF}/parallel.f90", ["transpose_xy", "transpose_yx", "transpose_yz", "transpose_zy"]))
    src.update(fx.load(ns, f"{REF}/mhdinit.f90", ["grid_initialize", "initial_calc_conserve_variable"]))
    src.update(fx.load(ns, f"{REF}/dealiasing.f90", ["dealias_initialize", "dealias"]))
    src.update(fx.load(ns, f"{REF}/AEBmod.f90", ["aeb_calc", "update_ksquare", "evolve_radius"]))
    src.update(fx.load(ns, f"{REF}/rktmod.f90", ["rkt_init", "rkt"]))
    src.update(fx.load(ns, f"{REF}/fftw.f90", ["from_xyz_to_zxy", "from_zxy_to_xyz", "transform_uu_real_to_fourier",
                                               "transform_uu_fourier_to_real"]))
    src.update(fx.load(ns, f"{REF}/mhdrhs.f90", ["calc_current_density_real", "calc_flux", "transform_flux_real_to_fourier",
                                                 "calc_rhs", "update_uu_prim_from_uu"]))
    src.update(fx.load(ns, f"{REF}/mhd.f90", ["evolve", "vardt", "calc_max_divb"]))
    src.update(fx.load(ns, f"{REF}/mhdrms.f90", ["calc_rms"]))
    return src


def load_reference_incompressible(ns):
    R = REFROOT + "/src_incompressible"
    src = {}
    src.update(fx.load(ns, f"{R}/parallel.f90", ["transpose_xy", "transpose_yx", "transpose_yz", "transpose_zy"]))
    src.update(fx.load(ns, f"{R}/mhdinit.f90", ["grid_initialize", "initial_calc_conserve_variable"]))
    src.update(fx.load(ns, f"{R}/dealiasing.f90", ["dealias_initialize", "dealias"]))
    src.update(fx.load(ns, f"{R}/AEBmod.f90", ["aeb_calc", "update_ksquare", "evolve_radius", "update_rho_p"]))
    src.update(fx.load(ns, f"{R}/rktmod.f90", ["rkt_init", "rkt"]))
    src.update(fx.load(ns, f"{R}/fftw.f90", ["from_xyz_to_zxy", "from_zxy_to_xyz", "transform_uu_real_to_fourier",
                                             "transform_uu_fourier_to_real"]))
    src.update(fx.load(ns, f"{R}/mhdrhs.f90", ["calc_current_density_real", "calc_gradient_velocity_real", "calc_flux_for_pressure",
                                               "transform_flux_for_pressure_real_to_fourier", "calc_pressure_fourier", "calc_flux",
                                               "transform_flux_real_to_fourier", "calc_rhs", "update_uu_prim_from_uu",
                                               "calc_divb_real", "calc_divv_real"]))
    src.update(fx.load(ns, f"{R}/mhd.f90", ["evolve", "vardt", "calc_max_divb", "calc_max_divv", "calc_max_divb_real", "calc_max_divv_real"]))
    return src


def run_case_incompressible(name, c, nsteps=2, pieces=True):
    """src_incompressible/mhd.f90:58-136,260-305 on one rank."""
    ns = build_namespace(c, incompressible=True)
    load_reference_incompressible(ns)
    st = ns["_storage"]
    out = {}
    prim = initial_primitive(c, seed=9)
    out["prim0"] = prim.copy()
    ns["grid_initialize"]()
    ns["dealias_initialize"]()
    if ns["if_aeb"]:
        ns["aeb_calc"](ns["radius"])
        ang = ns["corotating_angle"] if ns["if_corotating"] else 0.0
        ns["cos_cor_ang"], ns["sin_cor_ang"] = float(np.cos(ang)), float(np.sin(ang))
    else:
        ns["ur0"] = np.float64(0.0)                    # mhd.f90:88-90; tau_exp = r / Ur is then Inf as in IEEE Fortran, and unused
    st["uu"][...] = prim
    ns["initial_calc_conserve_variable"]()
    ns["transform_uu_real_to_fourier"]()
    out["uu_fourier0"] = st["uu_fourier"].copy()
    ns["vardt"]()
    out["dt0"] = ns["dt"]
    dts, times, radii, rho0s = [], [], [], []
    for istep in range(nsteps):
        if istep == 0 and pieces:                      # the pieces of the first stage, from the same text
            keep = {k: v.copy() for k, v in st.items()}
            ns["transform_uu_real_to_fourier"]()
            ns["calc_current_density_real"]()
            ns["calc_gradient_velocity_real"]()
            out["current_density_stage1"] = st["current_density"].copy()
            out["grad_velocity_stage1"] = st["grad_velocity"].copy()
            ns["calc_flux_for_pressure"]()
            out["flux_pressure_stage1"] = st["flux_pressure"].copy()
            ns["transform_flux_for_pressure_real_to_fourier"]()
            ns["calc_pressure_fourier"]()
            out["pressure_fourier_stage1"] = st["uu_fourier"][7].copy()
            ns["calc_flux"]()
            out["flux_stage1"] = st["flux"].copy()
            ns["transform_flux_real_to_fourier"]()
            ns["calc_rhs"]()
            out["fnl_stage1"] = st["fnl"].copy()
            for k, v in keep.items():
                st[k][...] = v
        ns["evolve"]()
        ns["time"] = ns["time"] + ns["dt"]
        ns["evolve_radius"](ns["time"])
        ns["vardt"]()
        dts.append(ns["dt"]); times.append(ns["time"]); radii.append(ns["radius"]); rho0s.append(ns["rho0"])
    out.update(uu=st["uu"].copy(), uu_prim=st["uu_prim"].copy(), uu_fourier=st["uu_fourier"].copy(),
               dt=np.array(dts), time=np.array(times), radius=np.array(radii), rho0=np.array(rho0s), p0=ns["p0"])
    ns["calc_max_divv"]()
    ns["calc_max_divb"]()
    out.update(max_divv=ns["max_divv"], max_divb=ns["max_divb"])
    ns["calc_divb_real"](); ns["calc_max_divb_real"]()
    out["max_divb_real"] = ns["max_divb"]
    ns["calc_divv_real"](); ns["calc_max_divv_real"]()
    out["max_divv_real"] = ns["max_divv"]
    # calc_rms is not run in this tree: src_incompressible/mhdrms.f90:74,84 read uu_prim(ix,iy,iz,4) while nvarPrim = 3
    # (mhdinit.f90:5) — out of bounds, undefined in the reference itself (the translator's bounds check stops there)
    out["switches"] = np.array([c[k] for k in sorted(c)], dtype=np.float64)
    out["switch_names"] = np.array(sorted(c))
    save_case(name, out, pieces)
    print(name, "dt", dts, "max_divV", out["max_divv"], "rho0", rho0s)


# src_compressible/2D: kz = 0, two passes, if_z_radial, square truncation (dealias option 3), if_limit_dt_increase,
# the user-defined external force
CASES_2D = {
    "c2d_hall_aeb_mask": dict(nx=32, ny=16, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1, if_resis=True,
                              if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False,
                              if_z_radial=False, if_limit_dt_increase=False, if_external_force=False),
    "c2d_zradial_square_explicit": dict(nx=16, ny=32, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=3,
                                        if_resis=True, if_resis_exp=True, if_visc=True, if_visc_exp=True, if_conserve_background=True,
                                        if_z_radial=True, if_limit_dt_increase=True, if_external_force=False),
    "c2d_external_force_filter": dict(nx=32, ny=32, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=2,
                                      if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False,
                                      if_z_radial=False, if_limit_dt_increase=False, if_external_force=True),
    # if_corotating in 2D: the oracle restates it, the library refuses it
    "c2d_corotating_oracle_only": dict(nx=16, ny=16, nz=1, if_hall=True, if_aeb=True, if_corotating=True, dealias_option=1,
                                       if_resis=True, if_resis_exp=True, if_visc=True, if_visc_exp=False, if_conserve_background=False,
                                       if_z_radial=False, if_limit_dt_increase=False, if_external_force=False),
    # 80-point x lines (5 * 16) and 48-point y lines (3 * 16)
    "c2d_lines80x48_hall_aeb_filter": dict(nx=80, ny=48, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=2,
                                           if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False,
                                           if_conserve_background=False, if_z_radial=False, if_limit_dt_increase=False,
                                           if_external_force=False),
}


def initial_primitive_2d(c, seed=3):
    import parity_common as pc
    _, prim = pc.make_case_2d(c["nx"], c["ny"], seed=seed)
    return prim


def run_case_2d(name, c, nsteps=3, pieces=True):
    """src_compressible/2D/mhd.f90 on one rank; vardt is called after every step here (the driver's dstep_calcdt
    cadence is host logic, not part of the path)."""
    R = REFROOT + "/src_compressible/2D"
    nx, ny = c["nx"], c["ny"]
    nxh = nx // 2 + 1
    ns = build_namespace(c)
    st = ns["_storage"]
    F = fx.FArray
    extra = dict(w_xy=np.zeros((1, ny, nxh), dtype=np.complex128), w_yx=np.zeros((1, ny, nxh), dtype=np.complex128),
                 external_force=np.zeros((1, 1, ny, nx)), external_force_fourier=np.zeros((1, 1, ny, nxh), dtype=np.complex128))
    for k, v in extra.items():
        st[k] = v
        ns[k] = F(v.T)
    ns.update(if_z_radial=c["if_z_radial"], if_limit_dt_increase=c["if_limit_dt_increase"], if_external_force=c["if_external_force"],
              nextern=1, isnanall=0, iproc=1)
    fx.load(ns, f"{R}/parallel.f90", ["transpose_xy", "transpose_yx"])
    fx.load(ns, f"{R}/mhdinit.f90", ["grid_initialize", "initial_calc_conserve_variable"])
    fx.load(ns, f"{R}/dealiasing.f90", ["dealias_initialize", "dealias"])
    fx.load(ns, f"{R}/AEBmod.f90", ["aeb_calc", "update_ksquare", "evolve_radius"])
    fx.load(ns, f"{R}/rktmod.f90", ["rkt_init", "rkt"])
    fx.load(ns, f"{R}/fftw.f90", ["transform_uu_real_to_fourier", "transform_uu_fourier_to_real"])
    fx.load(ns, f"{R}/mhdrhs.f90", ["calc_external_force_real", "calc_current_density_real", "calc_flux", "transform_flux_real_to_fourier",
                                    "calc_rhs", "update_uu_prim_from_uu"])
    fx.load(ns, f"{R}/mhd.f90", ["evolve", "vardt", "calc_max_divb", "checknan"])
    fx.load(ns, f"{R}/mhdrms.f90", ["calc_rms"])
    out = {}
    prim = initial_primitive_2d(c)
    out["prim0"] = prim.copy()
    ns["grid_initialize"]()
    ns["dealias_initialize"]()
    ns["aeb_calc"](ns["radius"])
    ang = ns["corotating_angle"] if ns["if_corotating"] else 0.0       # AEB_initialize (2D/AEBmod.f90:24-29)
    ns["cos_cor_ang"], ns["sin_cor_ang"] = float(np.cos(ang)), float(np.sin(ang))
    st["uu"][...] = prim
    ns["initial_calc_conserve_variable"]()
    ns["transform_uu_real_to_fourier"]()
    out["uu_fourier0"] = st["uu_fourier"].copy()
    ns["vardt"]()
    out["dt0"] = ns["dt"]
    dts, times, forces = [], [], []
    for istep in range(nsteps):
        if istep == 0 and pieces:
            keep = {k: v.copy() for k, v in st.items()}
            ns["calc_flux"]()
            out["flux_stage1"] = st["flux"].copy()
            out["expand_stage1"] = st["expand_term"].copy()
            ns["transform_flux_real_to_fourier"]()
            ns["calc_rhs"]()
            out["fnl_stage1"] = st["fnl"].copy()
            for k, v in keep.items():
                st[k][...] = v
        ns["evolve"]()
        forces.append(st["external_force"][0].copy())     # calc_external_force_real ran inside calc_flux with this step's time
        ns["time"] = ns["time"] + ns["dt"]
        ns["evolve_radius"](ns["time"])
        ns["vardt"]()
        dts.append(ns["dt"]); times.append(ns["time"])
    out.update(uu=st["uu"].copy(), uu_prim=st["uu_prim"].copy(), uu_fourier=st["uu_fourier"].copy(), k_square=st["k_square"].copy(),
               dt=np.array(dts), time=np.array(times), external_force=np.array(forces))
    ns["calc_max_divb"]()
    ns["calc_rms"]()
    ns["checknan"]()
    out.update(max_divb=ns["max_divb"], uu_ave=st["uu_ave"].copy(), uu_rms=st["uu_rms"].copy(), rho_u2=st["rho_u2"].copy(),
               isnanall=ns["isnanall"])
    out["switches"] = np.array([c[k] for k in sorted(c)], dtype=np.float64)
    out["switch_names"] = np.array(sorted(c))
    save_case(name, out, pieces)
    print(name, "dt", dts, "max_divB", ns["max_divb"], "isNanAll", ns["isnanall"])


# src_incompressible/2D
CASES_INCOMPRESSIBLE_2D = {
    "i2d_hall_aeb_mask": dict(nx=32, ny=16, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1, if_resis=True,
                              if_resis_exp=False, if_visc=True, if_visc_exp=False, if_conserve_background=False,
                              if_z_radial=False, if_limit_dt_increase=False, if_external_force=False),
    "i2d_square_explicit_limit": dict(nx=16, ny=32, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=3,
                                      if_resis=True, if_resis_exp=True, if_visc=True, if_visc_exp=True, if_conserve_background=True,
                                      if_z_radial=False, if_limit_dt_increase=True, if_external_force=False),
    # if_corotating (src_incompressible/2D/mhdrhs.f90:196-201,318-323; 2D/mhd.f90:602-607,653-658; 2D/AEBmod.f90:101-106)
    "i2d_corotating": dict(nx=16, ny=32, nz=1, if_hall=True, if_aeb=True, if_corotating=True, dealias_option=1, if_resis=True,
                           if_resis_exp=True, if_visc=True, if_visc_exp=False, if_conserve_background=False,
                           if_z_radial=False, if_limit_dt_increase=False, if_external_force=False),
    # 48-point x lines and 80-point y lines
    "i2d_lines48x80_hall_aeb_mask": dict(nx=48, ny=80, nz=1, if_hall=True, if_aeb=True, if_corotating=False, dealias_option=1,
                                         if_resis=True, if_resis_exp=False, if_visc=True, if_visc_exp=False,
                                         if_conserve_background=False, if_z_radial=False, if_limit_dt_increase=False,
                                         if_external_force=False),
}


def run_case_incompressible_2d(name, c, nsteps=3, pieces=False):
    """src_incompressible/2D/mhd.f90 on one rank (vardt after every step, as in run_case_2d)."""
    R = REFROOT + "/src_incompressible/2D"
    nx, ny = c["nx"], c["ny"]
    nxh = nx // 2 + 1
    ns = build_namespace(c, incompressible=True)
    st = ns["_storage"]
    F = fx.FArray
    extra = dict(w_xy=np.zeros((1, ny, nxh), dtype=np.complex128), w_yx=np.zeros((1, ny, nxh), dtype=np.complex128))
    for k, v in extra.items():
        st[k] = v
        ns[k] = F(v.T)
    ns.update(if_limit_dt_increase=c["if_limit_dt_increase"], isnanall=0, iproc=1, press0=1.0)   # &field press0 (2D/AEBmod.f90:122)
    fx.load(ns, f"{R}/parallel.f90", ["transpose_xy", "transpose_yx"])
    fx.load(ns, f"{R}/mhdinit.f90", ["grid_initialize", "initial_calc_conserve_variable"])
    fx.load(ns, f"{R}/dealiasing.f90", ["dealias_initialize", "dealias"])
    fx.load(ns, f"{R}/AEBmod.f90", ["aeb_calc", "update_ksquare", "evolve_radius", "update_rho_p"])
    fx.load(ns, f"{R}/rktmod.f90", ["rkt_init", "rkt"])
    fx.load(ns, f"{R}/fftw.f90", ["transform_uu_real_to_fourier", "transform_uu_fourier_to_real"])
    fx.load(ns, f"{R}/mhdrhs.f90", ["calc_current_density_real", "calc_gradient_velocity_real", "calc_flux_for_pressure",
                                    "transform_flux_for_pressure_real_to_fourier", "calc_pressure_fourier", "calc_flux",
                                    "transform_flux_real_to_fourier", "calc_rhs", "update_uu_prim_from_uu", "calc_divb_real", "calc_divv_real"])
    fx.load(ns, f"{R}/mhd.f90", ["evolve", "vardt", "calc_max_divb", "calc_max_divv", "calc_max_divb_real", "calc_max_divv_real", "checknan"])
    out = {}
    prim = initial_primitive_2d(c, seed=11)
    out["prim0"] = prim.copy()
    ns["grid_initialize"]()
    ns["dealias_initialize"]()
    ns["aeb_calc"](ns["radius"])
    ang = ns["corotating_angle"] if ns["if_corotating"] else 0.0       # AEB_initialize (src_incompressible/2D/AEBmod.f90:24-29)
    ns["cos_cor_ang"], ns["sin_cor_ang"] = float(np.cos(ang)), float(np.sin(ang))
    st["uu"][...] = prim
    ns["initial_calc_conserve_variable"]()
    ns["transform_uu_real_to_fourier"]()
    out["uu_fourier0"] = st["uu_fourier"].copy()
    ns["vardt"]()
    out["dt0"] = ns["dt"]
    dts, times, rho0s = [], [], []
    for istep in range(nsteps):
        ns["evolve"]()
        ns["time"] = ns["time"] + ns["dt"]
        ns["evolve_radius"](ns["time"])
        ns["vardt"]()
        dts.append(ns["dt"]); times.append(ns["time"]); rho0s.append(ns["rho0"])
    out.update(uu=st["uu"].copy(), uu_prim=st["uu_prim"].copy(), uu_fourier=st["uu_fourier"].copy(),
               dt=np.array(dts), time=np.array(times), rho0=np.array(rho0s))
    ns["calc_max_divv"]()
    out["max_divv"] = ns["max_divv"]
    ns["calc_divb_real"](); ns["calc_max_divb_real"]()
    out["max_divb_real"] = ns["max_divb"]
    ns["calc_divv_real"](); ns["calc_max_divv_real"]()
    out["max_divv_real"] = ns["max_divv"]
    ns["checknan"]()
    out["isnanall"] = ns["isnanall"]
    out["switches"] = np.array([c[k] for k in sorted(c)], dtype=np.float64)
    out["switch_names"] = np.array(sorted(c))
    save_case(name, out, pieces)
    print(name, "dt", dts, "max_divV", out["max_divv"], "rho0", rho0s)


def run_case_100_steps(name="hall_aeb_mask_100steps", every=10, nsteps=100):
    """The north star's second acceptance test on the executed reference source: 100 steps of the Principal loop (16 x 16 x 16, Hall +
    expanding box + spherical mask); mean energy density, mean cross helicity, max div B and the rms.dat row every 10 steps."""
    c = dict(CASES["hall_aeb_mask"], nx=16, ny=16, nz=16)
    ns = build_namespace(c)
    load_reference(ns)
    st = ns["_storage"]
    prim = initial_primitive(c, seed=21)
    out = {"prim0": prim.copy()}
    ns["grid_initialize"]()
    ns["dealias_initialize"]()
    ns["aeb_calc"](ns["radius"])
    st["uu"][...] = prim
    ns["initial_calc_conserve_variable"]()
    ns["transform_uu_real_to_fourier"]()
    ns["vardt"]()
    rows = []
    n = float(c["nx"] * c["ny"] * c["nz"])
    for istep in range(1, nsteps + 1):
        ns["evolve"]()
        ns["time"] = ns["time"] + ns["dt"]
        ns["evolve_radius"](ns["time"])
        ns["vardt"]()
        if istep % every == 0:
            ns["calc_max_divb"]()
            ns["calc_rms"]()
            uu, pr = st["uu"], st["uu_prim"]
            energy = uu[7].sum() / n                                           # mean total energy density uu(8)
            helicity = (pr[0] * uu[4] + pr[1] * uu[5] + pr[2] * uu[6]).sum() / n   # mean u.B
            rows.append(np.concatenate([[istep, ns["time"], ns["dt"], energy, helicity, ns["max_divb"]], st["uu_ave"], st["uu_rms"], st["rho_u2"]]))
    out.update(rows=np.array(rows), uu=st["uu"].copy(), switches=np.array([c[k] for k in sorted(c)], dtype=np.float64),
               switch_names=np.array(sorted(c)))
    np.savez_compressed(os.path.join(HERE, "ref_exec", name + ".npz"), **out)
    print(name, "time", ns["time"], "energy", rows[-1][3], "max_divB", rows[-1][5])


def run_initial_conditions(name="initial_conditions", nx=16, ny=8, nz=8):
    """The initial-condition hooks the benchmark and the stand-in driver rely on, from the reference's text
    (mhdinit.f90:183-260 background_fields_initialize case 3; :300-342 perturbation_initialize case 1; :695-823 case 7).
    The reference seeds a compiler-specific generator (random_seed(PUT = ir + 100), :705-727); as everywhere in this
    repository the three phase tables come from numpy.random.default_rng(ir + 100) instead — the hook below —
    and everything else (mode table, isotropy cut, polarisation along k x B0, amplitudes) is the Fortran's."""
    c = dict(CASES["hall_aeb_mask"], nx=nx, ny=ny, nz=nz)
    out = {}
    for ipert, extra in ((7, dict(nmodex=2, nmodey=2, nmodez=2, db0=0.1, dv0=0.1, drho0=0.01)),
                         (1, dict(db0=0.1, wave_number_jet=2))):
        ns = build_namespace(c)
        st = ns["_storage"]
        ns.update(rho0=1.0, t0=1.0, bx0=1.0, by0=0.25, bz0=-0.5, press0=1.0, ifield=3, ipert=ipert, nmode=0, correlation_vb=0.0,
                  b0=1.0, a0=0.05, **extra)
        seed = {}

        def random_seed(put=None, size=None, seed=seed):
            seed["s"] = int(put.a[0])

        def random_number(a, seed=seed):
            a.a[...] = np.random.default_rng(seed["s"]).random(a.a.size)

        ns.update(random_seed=random_seed, random_number=random_number, exit=lambda code: (_ for _ in ()).throw(SystemExit(code)))
        ns["_wrap_ints"]()
        fx.load(ns, f"{REF}/mhdinit.f90", ["grid_initialize"])
        ns["grid_initialize"]()
        tr = fx.Translator({k for k, v in ns.items() if isinstance(v, fx.FArray)})
        for sub, sel, val in (("background_fields_initialize", "ifield", "3"), ("perturbation_initialize", "ipert", str(ipert))):
            args, body = fx.subroutine_statements(f"{REF}/mhdinit.f90", sub)
            head = []
            for stmt in body:                                   # declarations and the statements before the select
                if stmt.startswith("select"):
                    break
                head.append(stmt)
            code = tr.subroutine(sub, args, head + fx.case_body(body, sel, val))
            exec(compile(code, f"<{sub}:case({val})>", "exec"), ns)
            ns[sub]()
        out[f"ipert{ipert}"] = st["uu"].copy()
    out["params"] = np.array([nx, ny, nz, 24.0, 12.0, 6.0, 1.0, 0.25, -0.5])
    np.savez_compressed(os.path.join(HERE, "ref_exec", name + ".npz"), **out)
    print(name, "rms B perturbation (ipert 7):", float(np.std(out["ipert7"][4])))


def run_parallel_start(nx, ny, nz, npe):
    """parallel_start + decompose_1d (parallel.f90:28-212,326-349) executed for every rank of a slab run
    (ndim_parallel = 1) on a fake MPI world: the decomposition tables and the MPI subarray types that define
    transpose_yz / transpose_zy.  Returns per rank: yj/zj offsets and sizes, and (full, sub, starts) of the send / receive
    types towards every peer."""
    ranks = []
    for ipe in range(npe):
        ns = fx.base_namespace()
        F = fx.FArray
        ints = lambda n: F(np.zeros(n, dtype=np.int64))  # noqa: E731
        objs = lambda n: F(np.empty(n, dtype=object))    # noqa: E731
        ns.update(dims=ints(2), periods=F(np.zeros(2, dtype=bool)), icoords=ints(2), reorder=True, ierr=0, ipe=0, npe=0, iproc=0, jproc=0,
                  myid_i=0, myid_j=0, ipe_cart=0, comm_cart=None, comm1d_i=None, comm1d_j=None,
                  xi_offset=ints(1), xi_size=ints(1), yi_offset=ints(1), yi_size=ints(1), yj_offset=ints(npe), yj_size=ints(npe),
                  zj_offset=ints(npe), zj_size=ints(npe), subarr_type_xy_send=objs(1), subarr_type_xy_recv=objs(1),
                  subarr_type_yz_send=objs(npe), subarr_type_yz_recv=objs(npe),
                  mpi_comm_world="world", mpi_real=4, mpi_complex=4, mpi_double_precision=8, mpi_double_complex=8, mpi_realtype=0,
                  mpi_complextype=0, mpi_order_fortran="F")

        def comm_rank(comm, me=ipe):
            if comm == "world":
                return fx.FInt(me)
            kind, color, key, members = comm
            return fx.FInt(sorted(members[color]).index(key))

        def comm_split(comm, color, key, me=ipe):
            # the members of every colour group, by key: in the 1 x npe grid of ndim_parallel = 1 the coordinates of rank r are (0, r)
            coords = [(0, r) for r in range(npe)]
            if int(color) == coords[me][1] and int(key) == coords[me][0]:      # split by j (comm1d_i): colour = my j
                members = {c[1]: [cc[0] for cc in coords if cc[1] == c[1]] for c in coords}
            else:                                                              # split by i (comm1d_j): colour = my i
                members = {c[0]: [cc[1] for cc in coords if cc[0] == c[0]] for c in coords}
            return ("sub", int(color), int(key), members)

        def cart_get(comm, nd, dims, periods, icoords, ierr, me=ipe):
            icoords.a[:] = (me // int(dims.a[1]), me % int(dims.a[1]))

        ns.update(mpi_init=lambda ierr: None, mpi_comm_rank=comm_rank, mpi_comm_size=lambda comm: fx.FInt(npe),
                  mpi_cart_create=lambda comm, nd, dims, periods, reorder: "cart", mpi_cart_get=cart_get,
                  mpi_cart_rank=lambda comm, icoords, me=ipe: fx.FInt(me), mpi_comm_split=comm_split,
                  mpi_type_create_subarray=lambda nd, full, sub, starts, order, typ: tuple([int(v) for v in x] for x in (full, sub, starts)),
                  mpi_type_commit=lambda t, ierr: None)
        fx.load(ns, f"{REF}/parallel.f90", ["decompose_1d", "parallel_start"])
        ns["parallel_start"](fx.FInt(nx), fx.FInt(ny), fx.FInt(nz), fx.FInt(8), fx.FInt(1))
        assert (int(ns["myid_i"]), int(ns["myid_j"]), int(ns["iproc"]), int(ns["jproc"])) == (0, ipe, 1, npe)
        ranks.append(dict(yj_offset=ns["yj_offset"].a.copy(), yj_size=ns["yj_size"].a.copy(), zj_offset=ns["zj_offset"].a.copy(),
                          zj_size=ns["zj_size"].a.copy(), xi_size=int(ns["xi_size"].a[0]),
                          yz_send=np.array([ns["subarr_type_yz_send"].a[q] for q in range(npe)], dtype=np.int64),
                          yz_recv=np.array([ns["subarr_type_yz_recv"].a[q] for q in range(npe)], dtype=np.int64)))
    return ranks


def make_parallel_fixtures():
    out = {}
    for nx, ny, nz, npe in [(16, 16, 16, 2), (16, 16, 16, 3), (16, 32, 16, 8), (32, 64, 32, 5), (16, 16, 16, 4),
                            (16, 48, 80, 5), (48, 80, 48, 3)]:       # odd-factor lengths: 48 = 9+9+9+9+12, 80 = 16 x 5 / 26+26+28
        for r, d in enumerate(run_parallel_start(nx, ny, nz, npe)):
            for k, v in d.items():
                out[f"{nx}x{ny}x{nz}_p{npe}_r{r}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_exec", "parallel_start.npz"), **out)
    print("parallel_start:", len(out), "arrays")


def save_case(name, out, pieces):
    """The initial spectrum and the k_square tables are oracle-only checks: kept with the stage pieces, dropped elsewhere."""
    if not pieces:
        for k in ("uu_fourier0", "k_square0", "k_square"):
            out.pop(k, None)
    np.savez_compressed(os.path.join(HERE, "ref_exec", name + ".npz"), **out)


def initial_primitive(c, seed=5):
    """Smooth O(1) primitive fields (rho, u, B, p) with content in every direction, [8, nz, ny, nx]."""
    import parity_common as pc
    from oracle import laps_oracle as lo
    p = lo.Params(nx=c["nx"], ny=c["ny"], nz=c["nz"], Lx=24.0, Ly=12.0, Lz=6.0)
    prim = lo.ic_uniform_background(p, bx0=1.0, press0=1.0)
    return lo.ic_turbulence(p, prim, 1.0, 0.0, 0.0, db0=0.1, dv0=0.1, drho0=0.01, nmodex=2, nmodey=2, nmodez=1,
                            seeds=(seed, seed + 15, seed + 31))


def run_case(name, c, nsteps=2, pieces=True):
    ns = build_namespace(c)
    load_reference(ns)
    st = ns["_storage"]
    out = {}
    prim = initial_primitive(c)
    out["prim0"] = prim.copy()
    # program mhd, mhd.f90:58-136 on one rank
    ns["grid_initialize"]()
    ns["dealias_initialize"]()
    if ns["if_aeb"]:                                   # AEB_initialize (AEBmod.f90:16-30) and mhd.f90:88-90
        ns["aeb_calc"](ns["radius"])
        ang = ns["corotating_angle"] if ns["if_corotating"] else 0.0
        ns["cos_cor_ang"], ns["sin_cor_ang"] = float(np.cos(ang)), float(np.sin(ang))
    else:
        ns["ur0"] = np.float64(0.0)                    # mhd.f90:88-90; tau_exp = r / Ur is then Inf as in IEEE Fortran, and unused
    st["uu"][...] = prim
    ns["initial_calc_conserve_variable"]()             # mhd.f90:121
    ns["transform_uu_real_to_fourier"]()               # :122
    out["k_square0"] = st["k_square"].copy()
    out["wave_numbers"] = np.concatenate([st["wave_number_x"], st["wave_number_y"], st["wave_number_z"]])
    out["uu_fourier0"] = st["uu_fourier"].copy()
    ns["vardt"]()                                      # :136
    out["dt0"] = ns["dt"]
    dts, times, radii = [], [], []
    for istep in range(nsteps):
        if istep == 0 and pieces:                      # the pieces of the first stage, from the same text
            keep = {k: v.copy() for k, v in st.items()}
            ns["calc_flux"]()
            out["flux_stage1"] = st["flux"].copy()
            out["expand_stage1"] = st["expand_term"].copy()
            out["current_density_stage1"] = st["current_density"].copy()
            ns["transform_flux_real_to_fourier"]()
            ns["calc_rhs"]()
            out["fnl_stage1"] = st["fnl"].copy()
            for k, v in keep.items():
                st[k][...] = v
        ns["evolve"]()                                 # mhd.f90:245
        ns["time"] = ns["time"] + ns["dt"]             # :246
        ns["evolve_radius"](ns["time"])                # :248
        ns["vardt"]()                                  # :285
        dts.append(ns["dt"]); times.append(ns["time"]); radii.append(ns["radius"])
    out.update(uu=st["uu"].copy(), uu_prim=st["uu_prim"].copy(), uu_fourier=st["uu_fourier"].copy(), k_square=st["k_square"].copy(),
               dt=np.array(dts), time=np.array(times), radius=np.array(radii))
    ns["calc_max_divb"]()
    ns["calc_rms"]()
    out.update(max_divb=ns["max_divb"], uu_ave=st["uu_ave"].copy(), uu_rms=st["uu_rms"].copy(), rho_u2=st["rho_u2"].copy())
    out["switches"] = np.array([c[k] for k in sorted(c)], dtype=np.float64)
    out["switch_names"] = np.array(sorted(c))
    os.makedirs(os.path.join(HERE, "ref_exec"), exist_ok=True)
    save_case(name, out, pieces)
    print(name, "dt", dts, "max_divB", ns["max_divb"])


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "ref_exec"), exist_ok=True)
    only = set(sys.argv[1:])                                   # case names to (re)make; none: everything
    if not only:
        make_parallel_fixtures()
        run_initial_conditions()
        run_case_100_steps()
    for i, (name, c) in enumerate(CASES.items()):              # stage pieces on the smaller grid of each tree (oracle-only check)
        if not only or name in only:
            run_case(name, c, pieces=(i == 1))
    for i, (name, c) in enumerate(CASES_INCOMPRESSIBLE.items()):
        if not only or name in only:
            run_case_incompressible(name, c, pieces=(i == 1))
    for i, (name, c) in enumerate(CASES_2D.items()):
        if not only or name in only:
            run_case_2d(name, c, pieces=(i in (0, 2)))     # the external-force case keeps its pieces (fnl(7) += force)
    for name, c in CASES_INCOMPRESSIBLE_2D.items():
        if not only or name in only:
            run_case_incompressible_2d(name, c)
