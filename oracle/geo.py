"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Geometric controller (ControllerType.GEO), one environment, numpy float64.  PARITY UNPINNED: the arithmetic lives in
sutra (``sutra_controller_geo``: init_proj_sparse / comp_dphi / comp_com), which the reference does not vendor; the
reference side that IS in the tree is
  init_controller_geo          shesha/init/rtc_init.py:418-448   pupil indices + mirrors handed to init_proj_sparse
  next_part_one_geo            shesha/supervisor/rlSupervisor.py:989-1013
                               target raytrace (atmosphere only) -> do_control -> apply_control -> raytrace through the mirrors
and the published algorithm is the least-squares projection

    com = -(IF IF^T)^-1 IF (phi - <phi>)       IF [nactu][pupil points], phi over the same points.

This restatement is deliberately independent of the product's tables: the influence rows come from the oracle's own
mirror model (aoframe.pzt_shape / tt_shape through loop.dm_phase, one unit command at a time), and the fit is
``numpy.linalg.lstsq`` / LSQR on the phase itself, not a precomputed inverse.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from . import loop


def influence_matrix(tab):
    """Sparse [pupil points][nactu]: column k = phase of a unit command on actuator k over the pupil (microns)."""
    m = np.asarray(tab["mpupil"]) != 0
    idx = np.nonzero(m.ravel())[0]
    nactu = int(tab["nactu"])
    cols = []
    for k in range(nactu):
        v = np.zeros(nactu, np.float32)
        v[k] = 1.0
        ph = loop.dm_phase(tab, v).astype(np.float64).ravel()[idx]
        cols.append(sp.csc_matrix(ph[:, None]))
    return sp.hstack(cols).tocsr(), idx


def geo_command(IFt, idx, phase):
    """Least-squares mirror command cancelling `phase` [n][n] over the pupil; returns (com, residual over the pupil)."""
    phi = np.asarray(phase, np.float64).ravel()[idx]
    phi = phi - phi.mean()
    com = sla.lsqr(IFt, -phi, atol=1e-13, btol=1e-13, iter_lim=20000)[0]
    return com, phi + IFt @ com
