"""Aero-table loading — mirror of `Aerodynamics.load_aerodata` / `rescale_aerodata`
(reference aerodynamics.jl:11-36).  The force model itself (aero_force, aerodynamics.jl:38-58) runs
on the device; this module only brings the tables to the boundary in the reference's layout.
"""
from __future__ import annotations

import os

import numpy as np

from .defns import AeroTable, AtmosphericData, ExoatmosphericData

# aerodynamics.jl:17-18 —  mach = 0.0:0.025:1.5 ; aoa = cosd(180):1/90:cosd(0)
N_COS, N_MACH = 181, 61
COS0, DCOS = -1.0, 1.0 / 90.0
MACH0, DMACH = 0.0, 0.025


def _table(col: np.ndarray) -> AeroTable:
    # reshape(col, length(aoa), length(mach)) is column-major: cos(aoa) varies fastest (aerodynamics.jl:19)
    return AeroTable(np.asfortranarray(np.asarray(col, dtype=np.float64).reshape((N_COS, N_MACH), order="F")),
                     COS0, DCOS, MACH0, DMACH)


def load_aerodata(liftdrag, finforce=None) -> AtmosphericData:
    """aerodynamics.jl:11-28.  Accepts the reference CSV (`aoa,mach,drag,lift,torque`, numbers in
    newtons) or an .npz with arrays `drag`, `lift`, `torque` of 181*61 samples in CSV row order.
    `finforce` is accepted and ignored, as in the reference (aerodynamics.jl:23-26 reads and drops it)."""
    path = os.fspath(liftdrag)
    if path.endswith(".npz"):
        z = np.load(path)
        drag, lift, torque = z["drag"], z["lift"], z["torque"]
    else:
        raw = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64)
        with open(path) as fh:
            header = fh.readline().strip().split(",")
        drag, lift, torque = (raw[:, header.index(n)] for n in ("drag", "lift", "torque"))
    if drag.size != N_COS * N_MACH:
        raise ValueError(f"aero table must hold {N_COS}x{N_MACH} samples, got {drag.size}")
    return AtmosphericData(_table(drag), _table(lift), _table(torque), 1.0, 1.0)


def rescale_aerodata(data, Ul: float, Ut: float, Um: float):
    """aerodynamics.jl:30-36."""
    if isinstance(data, ExoatmosphericData):
        return data
    return AtmosphericData(data.drag_itrp, data.lift_itrp, data.trq_itrp, 1 / (Ul * Um / Ut ** 2), 1 / Ul)


# aero/AeroTable.jl:94-112 — fin.csv: header `lift,drag,mach,aoa`; rows = 901 deflection angles (0:0.1:90 deg) x 60 Mach
# numbers (0.01:0.025:1.485, fastest varying).  Read and dropped by the reference (aerodynamics.jl:23-26).
FIN_N_MACH, FIN_N_DEFL = 60, 901
FIN_MACH0, FIN_DMACH = 0.01, 0.025
FIN_DEFL0, FIN_DDEFL = 0.0, 0.1


def load_fin_table(finforce, n_mach: int = FIN_N_MACH, n_defl: int = FIN_N_DEFL):
    """The fin-force table of `aero/fin.csv` as two (n_mach, n_defl) column-major arrays (lift, drag): the samples
    `scvx_set_fin_table` stages (no consumer in the reference).  The axes are read from the file's own
    `mach` / `aoa` columns."""
    raw = np.loadtxt(os.fspath(finforce), delimiter=",", skiprows=1, dtype=np.float64)
    with open(os.fspath(finforce)) as fh:
        header = fh.readline().strip().split(",")
    if raw.shape[0] != n_mach * n_defl:
        raise ValueError(f"fin table must hold {n_mach}x{n_defl} rows, got {raw.shape[0]}")
    col = {n: raw[:, header.index(n)] for n in ("lift", "drag", "mach", "aoa")}
    mach = col["mach"].reshape((n_mach, n_defl), order="F")[:, 0]
    defl = col["aoa"].reshape((n_mach, n_defl), order="F")[0, :]
    axes = (float(mach[0]), float(mach[1] - mach[0]), float(defl[0]), float(defl[1] - defl[0]))
    lift, drag = (np.asfortranarray(col[n].reshape((n_mach, n_defl), order="F")) for n in ("lift", "drag"))
    return lift, drag, axes
