"""Synthetic cases of BASELINE.json's configs (host side, numpy): initial
macroscopic fields in lattice units on the barycentres of a LevelDesc.

  cfg1  D3Q19 BGK Taylor-Green vortex, periodic cube      (TGV_Simple_Re800/musubi.lua:36-60 values)
  cfg2  D3Q19 TRT lid-driven cavity, bounce-back walls + velocity_bounceback lid
  cfg3  D3Q27 MRT periodic "channel": TGV-like vortex + uniform mean flow

The equilibrium used to fill the PDFs is the generic second-order polynomial
w_i rho (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 u^2) (mus_flow_module.fpp:484-589 with
zero strain rate); parity tests upload the oracle's own initial state instead.
"""
import math

import numpy as np

_CX27 = np.array([
    [-1, 0, 0], [0, -1, 0], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1],
    [0, -1, -1], [0, -1, 1], [0, 1, -1], [0, 1, 1],
    [-1, 0, -1], [1, 0, -1], [-1, 0, 1], [1, 0, 1],
    [-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0],
    [-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1],
    [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=np.float64)


def stencil(QQ):
    cx = np.vstack([_CX27[:QQ - 1], np.zeros((1, 3))])
    n = (cx ** 2).sum(axis=1)
    if QQ == 19:
        w = np.where(n == 0, 1.0 / 3.0, np.where(n == 1, 1.0 / 18.0, 1.0 / 36.0))
    else:
        w = np.where(n == 0, 8.0 / 27.0, np.where(n == 1, 2.0 / 27.0, np.where(n == 2, 1.0 / 54.0, 1.0 / 216.0)))
    return cx, w


def equilibrium_state(QQ, rho, vel, nSize, chunk=1 << 20):
    """AOS state array (nSize*QQ) with f = fEq(rho, vel) for the first len(rho) elements."""
    cx, w = stencil(QQ)
    n = rho.shape[0]
    out = np.zeros(nSize * QQ)
    view = out[:n * QQ].reshape(n, QQ)
    for s in range(0, n, chunk):
        r = rho[s:s + chunk, None]
        u = vel[s:s + chunk]
        cu = u @ cx.T
        usq = (u * u).sum(axis=1)[:, None]
        view[s:s + chunk] = w[None, :] * r * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * usq)
    return out


def taylor_green(ld, u0=0.09 / math.sqrt(3.0), mean=(0.0, 0.0, 0.0)):
    """rho, vel (lattice units) of the Taylor-Green vortex on a 2*pi periodic cube."""
    x = ld.barycenters((0.0, 0.0, 0.0), 2.0 * math.pi)
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    vel = np.stack([u0 * np.sin(X) * np.cos(Y) * np.cos(Z) + mean[0],
                    -u0 * np.cos(X) * np.sin(Y) * np.cos(Z) + mean[1],
                    np.zeros_like(X) + mean[2]], axis=1)
    p = u0 * u0 / 16.0 * (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2.0)
    rho = 1.0 + 3.0 * p
    return rho, vel


def cavity_rest(ld):
    return np.ones(ld.nElems), np.zeros((ld.nElems, 3))


def lid_values(ld, u_lid=(0.05, 0.0, 0.0)):
    """per-link lattice velocity of the 'lid' boundary (id 2), constant in time."""
    for bc in ld.bc:
        if bc["id"] == 2:
            return np.tile(np.asarray(u_lid, dtype=np.float64), (len(bc["links"]), 1))
    return np.zeros((0, 3))
