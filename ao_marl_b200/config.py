"""Parameter holders and loaders for the AO simulation.

Mirrors the *schema* of the reference's COMPASS parameter classes
(/root/reference/shesha/config/P{LOOP,GEOM,TEL,ATMOS,TARGET,WFS,DMS,CENTROIDER,CONTROLLER}.py):
same field names, same defaults, and the same ``set_<field>(value)`` / ``get_<field>()`` accessors,
so a parameter file written for the reference (``import shesha.config as conf``) loads unchanged
through :func:`load_config_from_file`.  Derived fields filled in by the init builders keep the
reference's underscore names (``_pdiam``, ``_Nfft``, ``_n1`` ...).
"""
import importlib.util
import os
import sys
import types

import numpy as np


class _Params:
    """Generic holder: attributes + set_x/get_x accessors (reference style)."""

    _defaults = {}
    _int_arrays = ()
    _float_arrays = ()

    def __init__(self, **kw):
        for k, v in self._defaults.items():
            object.__setattr__(self, k, v.copy() if isinstance(v, np.ndarray) else v)
        for k, v in kw.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if v is not None:
            if k in self._int_arrays:
                v = np.atleast_1d(np.asarray(v, dtype=np.int32))
            elif k in self._float_arrays:
                v = np.atleast_1d(np.asarray(v, dtype=np.float32))
        object.__setattr__(self, k, v)

    def __getattr__(self, name):
        if name.startswith("set_"):
            field = name[4:]
            return lambda v: setattr(self, field, v)
        if name.startswith("get_"):
            field = name[4:]
            return lambda: getattr(self, field)
        raise AttributeError(name)


class Param_loop(_Params):
    _defaults = dict(niter=0, ittime=0.0, devices=np.array([0], dtype=np.int32))


class Param_geom(_Params):
    _defaults = dict(is_init=False, ssize=0, zenithangle=0.0, apod=False, pupdiam=0, cent=0.0,
                     _pixsize=0.0, _ipupil=None, _mpupil=None, _spupil=None, _apodizer=None,
                     _p1=0, _p2=0, _n=0, _n1=0, _n2=0, _phase_ab_M1=None, _phase_ab_M1_m=None)


class Param_tel(_Params):
    _defaults = dict(diam=0.0, cobs=0.0, type_ap="Generic", t_spiders=-1.0, spiders_type=None,
                     pupangle=0.0, nbrmissing=0, gap=0.0, referr=0.0, std_piston=0.0, std_tt=0.0)


class Param_atmos(_Params):
    _defaults = dict(nscreens=0, r0=None, pupixsize=None, L0=None, dim_screens=None, alt=None,
                     winddir=None, windspeed=None, frac=None, _deltax=None, _deltay=None,
                     seeds=None)
    _float_arrays = ("L0", "alt", "winddir", "windspeed", "frac")


class Param_target(_Params):
    _defaults = dict(apod=False, Lambda=None, xpos=None, ypos=None, mag=None, zerop=1.0,
                     dms_seen=None)
    _int_arrays = ("dms_seen",)


class Param_wfs(_Params):
    _defaults = dict(type=None, nxsub=0, npix=0, pixsize=0.0, Lambda=0.0, optthroughput=0.0,
                     fracsub=0.0, open_loop=False, atmos_seen=0, dms_seen=None, roket=False,
                     is_low_order=False, xpos=0.0, ypos=0.0, gsalt=0.0, gsmag=0.0, zerop=0.0,
                     noise=0.0, kernel=0.0, nphotons4imat=1.0e5, nPupils=0,
                     _pdiam=0, _Nfft=0, _Ntot=0, _nrebin=0, _nvalid=0, _nphotons=0.0, _subapd=0.0,
                     _fluxPerSub=None, _qpixsize=0.0, _validpuppixx=None, _validpuppixy=None,
                     _validsubsx=None, _validsubsy=None, _isvalid=None, _phasemap=None,
                     _hrmap=None, _binmap=None, _halfxy=None, _ftkernel=None)
    _int_arrays = ("dms_seen",)

    def __init__(self, roket=False, **kw):
        super().__init__(roket=roket, **kw)


class Param_dm(_Params):
    _defaults = dict(type=None, nact=0, alt=0.0, thresh=0.0, coupling=0.2, gain=1.0,
                     pupoffset=None, unitpervolt=0.01, push4imat=1.0, margin_out=None,
                     margin_in=0.0, pzt_extent=5.0, type_pattern=None, influ_type="default",
                     _pitch=None, _ntotact=None, _influsize=None, _n1=None, _n2=None,
                     _influ=None, _xpos=None, _ypos=None, _i1=None, _j1=None, _influpos=None,
                     _ninflu=None, _influstart=None, _dim_screen=0)


class Param_centroider(_Params):
    _defaults = dict(nwfs=None, type=None, nslope=0, type_fct="gauss", weights=None, nmax=10,
                     thresh=1.0e-4, width=0.0, sizex=None, sizey=None, interpmat=None,
                     method=1, pyrscale=0, filter_TT=False, _nslope=0)


class Param_controller(_Params):
    _defaults = dict(type=None, nwfs=None, nvalid=0, nslope=0, ndm=None, nactu=0, _imat=None,
                     _cmat=None, maxcond=None, TTcond=None, delay=None, gain=None, nkl=None,
                     modopti=False, nrec=2048, nmodes=None, gmin=0.0, gmax=1.0, ngain=15,
                     do_kl_imat=False, klpush=None, klgain=None, nstates=0)
    _int_arrays = ("nwfs", "ndm")


_SHIM_NAMES = ("Param_loop", "Param_geom", "Param_tel", "Param_atmos", "Param_target",
               "Param_wfs", "Param_dm", "Param_centroider", "Param_controller")


def _shim_module():
    m = types.ModuleType("shesha.config")
    for n in _SHIM_NAMES:
        setattr(m, n, globals()[n])
    return m


_CFG_FIELDS = ("p_loop", "p_geom", "p_tel", "p_atmos", "p_dms", "p_targets", "p_wfss",
               "p_centroiders", "p_controllers", "simul_name")


def _finish(ns):
    for f in _CFG_FIELDS:
        if not hasattr(ns, f):
            setattr(ns, f, None)
    return ns


def load_config_from_file(filename_path: str):
    """Load a parameter set (same call as shesha.util.utilities.load_config_from_file, utilities.py:159-185).

    Accepts (1) the name of a built-in production set (``"production_sh_10x10_2m.py"``, with or without
    ``.py`` / leading directories that do not exist on disk) or (2) a path to a reference-format
    parameter file, executed with ``shesha.config`` bound to this module's classes.
    """
    from . import params
    base = os.path.basename(filename_path)
    name = base[:-3] if base.endswith(".py") else base
    if os.path.isfile(filename_path):
        saved = {k: sys.modules.get(k) for k in ("shesha", "shesha.config")}
        shesha = types.ModuleType("shesha")
        shesha.config = _shim_module()
        sys.modules["shesha"] = shesha
        sys.modules["shesha.config"] = shesha.config
        try:
            spec = importlib.util.spec_from_file_location("aomarl_par_" + name, filename_path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        return _finish(mod)
    if name in params.REGISTRY:
        return _finish(params.REGISTRY[name]())
    raise ValueError("Config file must be an existing .py file or one of %s" %
                     sorted(params.REGISTRY))
