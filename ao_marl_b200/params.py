"""Built-in production parameter sets.

One generator covers the family of parameter files the reference ships under
/root/reference/data/par/par4rl/production/ (values cited from those files: e.g.
production_sh_10x10_2m.py:6-169, production_sh_40x40_8m_3layers.py:6-173,
production_sh_40x40_8m_3layers_d0_noise.py:15-148).  Three layouts exist there:

* ``geo``   : two WFS / two targets / four DMs ordered [pzt, pzt_geo, tt, tt_geo], an LS controller
              on DMs [0, 2] + a GEO controller on DMs [1, 3]         (10x10_2m, 40x40_8m_3layers*, *_v_*, *_dir_*)
* ``noise`` : a second noise-free WFS, two DMs [pzt, tt], one LS controller   (*_d0_noise, *_d1_noise)
* ``roket`` : single WFS (roket flag), two DMs, one LS controller             (*_same_dir_roket)

The WFS list matters: with two equal-nxsub sensors the reference initialises WFS 1 first
(geom_init.py:79-94), which is what gives WFS 0 its Nfft = 64 / nrebin = 2 geometry.
"""
import types

import numpy as np

from . import config as conf


def production(simul_name, *, nxsub, diam, layout="geo", delay=1.0, noise=-1.0, gain=0.7,
               gsmag=4.0, r0=0.16, frac=(1.0,), alt=(0.0,), windspeed=(20.0,), winddir=(45.0,),
               L0=(1.0e5,)):
    ns = types.SimpleNamespace(simul_name=simul_name)
    ns.p_loop = conf.Param_loop(niter=2000, ittime=0.002)
    ns.p_geom = conf.Param_geom(zenithangle=0.0)
    ns.p_tel = conf.Param_tel(diam=float(diam), cobs=0.12)
    ns.p_atmos = conf.Param_atmos(r0=float(r0), nscreens=len(frac), frac=list(frac),
                                  alt=list(alt), windspeed=list(windspeed),
                                  winddir=list(winddir), L0=list(L0))
    geo = layout == "geo"
    main_dms = [0, 2] if geo else [0, 1]

    def target(dms_seen):
        return conf.Param_target(dms_seen=dms_seen, xpos=0.0, ypos=0.0, Lambda=1.65, mag=10.0)

    ns.p_targets = [target(main_dms)] + ([target([1, 3])] if geo else [])

    def wfs(dms_seen, noise_, roket=False):
        return conf.Param_wfs(roket=roket, type="sh", nxsub=nxsub, npix=16, dms_seen=dms_seen,
                              pixsize=0.25, fracsub=0.8, xpos=0.0, ypos=0.0, Lambda=0.5,
                              gsmag=float(gsmag), optthroughput=0.12, zerop=1.0e11,
                              noise=float(noise_), atmos_seen=1)

    if layout == "geo":
        ns.p_wfss = [wfs(main_dms, noise), wfs([1, 3], -1.0)]
    elif layout == "noise":
        ns.p_wfss = [wfs(main_dms, noise), wfs(main_dms, -1.0)]
    elif layout == "roket":
        ns.p_wfss = [wfs(main_dms, noise, roket=True)]
    else:
        raise ValueError(layout)

    def pzt():
        return conf.Param_dm(type="pzt", nact=nxsub + 1, alt=0.0, thresh=0.3, coupling=0.2,
                             unitpervolt=0.01, push4imat=100.0)

    def tt():
        return conf.Param_dm(type="tt", alt=0.0, unitpervolt=0.0005, push4imat=10.0)

    ns.p_dms = [pzt(), pzt(), tt(), tt()] if geo else [pzt(), tt()]
    ns.p_centroiders = [conf.Param_centroider(nwfs=0, type="cog")]
    ns.p_controllers = [conf.Param_controller(type="ls", nwfs=[0], ndm=main_dms, maxcond=1500.0,
                                              delay=float(delay), gain=float(gain))]
    if geo:
        ns.p_centroiders.append(conf.Param_centroider(nwfs=1, type="cog"))
        ns.p_controllers.append(conf.Param_controller(type="geo", nwfs=[1], ndm=[1, 3],
                                                      maxcond=1500.0, delay=0.0,
                                                      gain=float(gain)))
    return ns


_L3 = dict(nxsub=40, diam=8.0, frac=(0.6, 0.25, 0.15), alt=(0.0, 4500.0, 14000.0),
           L0=(1.0e5, 1.0e5, 1.0e5))


def _mk(name, **kw):
    return lambda: production(name, **kw)


REGISTRY = {}


def _register(name, **kw):
    REGISTRY[name] = _mk(name, **kw)


_register("production_sh_10x10_2m", nxsub=10, diam=2.0)
for _dirname, _dirs in (("", (0, 45, 90)), ("_dir_0_15_30", (0, 15, 30)), ("_same_dir", (0, 0, 0))):
    for _vname, _v, _g in (("", (15, 10, 20), 0.7), ("_v_10_5_15", (10, 5, 15), 0.6),
                           ("_v_20_15_25", (20, 15, 25), 0.7)):
        _register("production_sh_40x40_8m_3layers%s%s" % (_dirname, _vname),
                  windspeed=_v, winddir=_dirs, gain=_g, **_L3)
_register("production_sh_40x40_8m_3layers_d0_noise", layout="noise", delay=0.0, noise=3.0,
          gain=0.3, gsmag=9.0, windspeed=(15, 10, 20), winddir=(0, 45, 90), **_L3)
_register("production_sh_40x40_8m_3layers_d1_noise", layout="noise", delay=1.0, noise=3.0,
          gain=0.65, gsmag=9.0, windspeed=(15, 10, 20), winddir=(0, 45, 90), **_L3)
_register("production_sh_40x40_8m_3layers_same_dir_roket", layout="roket",
          windspeed=(15, 10, 20), winddir=(0, 0, 0), **_L3)
