"""shared set-up of parity cases: the oracle scheme and the device scheme fed
with identical inputs."""
import numpy as np


def make_pair(mb, mo, level, ident, omega, kind="periodic", rank=0, nranks=1, lambda_=0.25,
              omega_bulk=None, ic="tgv", u_lid=(0.05, 0.02, 0.0), outlet="pressure_expol",
              rho_out=1.0):
    from musubi_b200 import cases
    QQ = 19 if ident["layout"] == "d3q19" else 27
    ld = mb.LevelDesc(level, QQ, kind, rank, nranks)
    old = mo.build_level_desc(level, QQ, kind, rank, nranks)
    ob = omega if omega_bulk is None else omega_bulk
    ref = mo.Scheme(old, ident["relaxation"], ident["kind"], omega=omega, lambda_=lambda_, omega_bulk=ob)
    if ic == "tgv":
        rho, vel = cases.taylor_green(ld, mean=(0.01, -0.02, 0.015))
    else:
        rho, vel = cases.cavity_rest(ld)
    ref.init_equilibrium(rho, vel)
    # the reference derives omega from the lattice viscosity every step
    # (mus_update_relaxParamKine): hand the device exactly that value
    omega_eff = float(1.0 / (3.0 * ref.visc[0] + 0.5))
    bc_kind = {3: outlet} if kind == "channel" else None
    sch = mb.Scheme(ident, ld, omega_eff, lambda_=lambda_, omega_bulk=ob, bc_kind=bc_kind)
    sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
    if kind == "channel":
        # inlet (id 2): velocity_bounceback with u_lid as the inflow velocity; outlet (id 3): the
        # chosen pressure boundary at lattice density rho_out
        v = cases.lid_values(ld, u_lid)
        ref.bc_vel[2] = v
        sch.set_bc_values(level, 2, v)
        nOut = len([b for b in ld.bc if b["id"] == 3][0]["elems"])
        ref.bc_kind[3] = outlet
        ref.bc_rho[3] = np.full(nOut, rho_out)
        sch.set_bc_values(level, 3, ref.bc_rho[3])
        # pressure_expol reads the auxField of the previous step: hand over the initial one
        from musubi_b200._lib import check, lib
        check(lib.musb200_aux_upload(level, ref.aux.ctypes.data))
    if kind == "cavity":
        v = cases.lid_values(ld, u_lid)
        ref.bc_vel[2] = v
        sch.set_bc_values(level, 2, v)
    return ld, old, ref, sch


def rel_diff(got, exp):
    den = np.maximum(np.abs(exp), 1e-300)
    return float(np.max(np.abs(got - exp) / den))
