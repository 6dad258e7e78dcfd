"""Assembly of all static tables of one AO configuration, in the reference's initialisation order
(GenericSupervisor._init_components, shesha/supervisor/genericSupervisor.py:116-142:
telescope -> atmosphere -> DMs -> targets -> WFS -> RTC), restricted to what controller 0 drives.

Everything here is host-side numpy run once; the results are uploaded to the GPU by
`ao_marl_b200.lib.Simulator`.  The interaction matrix itself is measured on the GPU (it needs
the Shack-Hartmann pipeline) by `ao_marl_b200.calibration`.
"""
import numpy as np

from .init import atmos as atmos_b
from .init import dm as dm_b
from .init import geom as geom_b
from .init import rtc as rtc_b


class StaticTables:
    """Plain container; see `build_static` for the fields."""

    def as_oracle_dict(self):
        """Tables in the dict layout the numpy oracle (oracle/aoframe.py, oracle/loop.py) consumes."""
        w = self.p_wfs
        return dict(
            n=self.n, mpupil=self.mpupil, spupil=self.spupil,
            nscreens=self.nscreens, dim_screens=self.dim_screens, deltax=self.deltax,
            deltay=self.deltay, r0_layers=self.r0_layers, A=self.A, B=self.B,
            istx_pos=self.istx, wfs_xoff=self.wfs_xoff, wfs_yoff=self.wfs_yoff,
            wfs=dict(nvalid=w._nvalid, pdiam=w._pdiam, Nfft=w._Nfft, Ntot=w._Ntot, npix=w.npix,
                     nxsub=w.nxsub, nrebin=w._nrebin, Lambda=w.Lambda, phasemap=w._phasemap,
                     binmap=w._binmap, halfxy=w._halfxy, fluxPerSub=w._fluxPerSub_list,
                     nphotons=w._nphotons, pixsize=w.pixsize, validsubsx=w._validsubsx,
                     validsubsy=w._validsubsy, noise=w.noise, tile_origin=w._tile_origin),
            pzt=dict(influ=self.p_pzt._influ, i1=self.p_pzt._i1, j1=self.p_pzt._j1,
                     dim=self.pzt_dim, off=self.pzt_off, nact=self.p_pzt._ntotact),
            tt=dict(influ=self.p_tt._influ, dim=self.tt_dim, off=self.tt_off),
            nactu=self.nactu, nslopes=self.nslopes, gain=self.gain, delay=self.delay)


def build_static(config, verbose=False):
    """Run the host builders on a loaded parameter set; returns StaticTables.

    Mutates the parameter objects exactly like the reference's init functions do (derived fields
    are stored on them under the same underscore names)."""
    t = StaticTables()
    t.config = config
    r0, ittime = config.p_atmos.r0, config.p_loop.ittime
    geom_b.tel_init(config.p_geom, config.p_tel, r0, ittime, config.p_wfss)
    atmos_b.atmos_init(config.p_atmos, config.p_tel, config.p_geom, ittime, config.p_wfss,
                       config.p_targets)
    dm_b.dm_init(config.p_dms, config.p_tel, config.p_geom, config.p_wfss)

    ctrl = config.p_controllers[0]
    if ctrl.type != "ls":
        raise NotImplementedError("controller 0 must be the least-squares integrator")
    if len(ctrl.nwfs) != 1 or len(ctrl.ndm) != 2:
        raise NotImplementedError("hot path = one SH sensor driving one piezo + one tip-tilt mirror")
    t.p_wfs = config.p_wfss[int(ctrl.nwfs[0])]
    t.wfs_index = int(ctrl.nwfs[0])
    dms = [config.p_dms[int(i)] for i in ctrl.ndm]
    if [d.type for d in dms] != ["pzt", "tt"]:
        raise NotImplementedError("controller 0 must drive [pzt, tt]")
    t.p_pzt, t.p_tt = dms
    if not dm_b.stamp_is_shared(t.p_pzt):
        raise NotImplementedError("actuator stamps differ: the shared-stamp kernels do not apply")

    # unseen-actuator filtering (rtc_init.py:116-123)
    t.imat_geom = rtc_b.imat_geom(t.p_wfs, dms, config.p_geom)
    rtc_b.correct_dm(dms, config.p_geom, t.imat_geom)

    g = config.p_geom
    t.n = int(g._n)
    t.pupdiam = int(g.pupdiam)
    t.mpupil = g._mpupil.astype(np.float32)
    t.spupil = g._spupil.astype(np.float32)
    a = config.p_atmos
    t.nscreens = int(a.nscreens)
    t.dim_screens = np.asarray(a.dim_screens, dtype=np.int64)
    t.deltax = np.asarray(a._deltax, dtype=np.float64)
    t.deltay = np.asarray(a._deltay, dtype=np.float64)
    t.r0_layers = np.asarray(a._r0_layers, dtype=np.float64)
    ops = atmos_b.layer_operands(a)
    t.A = [o[0] for o in ops]
    t.B = [o[1] for o in ops]
    t.istx = [o[2] for o in ops]
    offs = [rtc_b.atmos_offset_in_wfs(a, g, t.p_wfs, l) for l in range(t.nscreens)]
    t.wfs_xoff = np.array([o[0] for o in offs])
    t.wfs_yoff = np.array([o[1] for o in offs])
    t.pzt_dim = int(max(t.p_pzt._n2 - t.p_pzt._n1 + 1, t.n))
    t.tt_dim = int(max(t.p_tt._n2 - t.p_tt._n1 + 1, t.n))
    t.pzt_off = rtc_b.dm_offset_in_wfs(t.p_pzt, g)
    t.tt_off = rtc_b.dm_offset_in_wfs(t.p_tt, g)
    t.nactu = int(t.p_pzt._ntotact + t.p_tt._ntotact)
    t.nslopes = 2 * int(t.p_wfs._nvalid)
    t.cog_offset, t.cog_scale = rtc_b.centroider_constants(t.p_wfs)
    t.gain = float(ctrl.gain)
    t.delay = float(ctrl.delay)
    ctrl.nactu, ctrl.nslope, ctrl.nvalid = t.nactu, t.nslopes, int(t.p_wfs._nvalid)
    return t


def build_basis(t):
    """Btt / P from the influence functions (rlSupervisor.py:169-176 -> modalBasis -> basis.compute_btt)."""
    IF = rtc_b.influence_matrix([t.p_pzt, t.p_tt], t.config.p_geom)
    n = IF.shape[1]
    IFtt = IF[:, n - 2:].toarray()
    IFpzt = IF[:, :n - 2]
    t.Btt, t.P = rtc_b.compute_btt(IFpzt, IFtt)
    return t.Btt, t.P
