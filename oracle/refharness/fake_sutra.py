"""TEST INFRASTRUCTURE -- recording stand-ins for the un-vendored COMPASS modules.

The reference imports ``carmaWrap.context`` and ``sutraWrap.{Atmos,Telescope,Dms,Sensors,
Target,Rtc_FFF,...}`` (shesha/sutra_wrap.py:46-72).  These fakes let the reference's OWN
init code (tel_init, atmos_init, dm_init, wfs_init, target_init, imat_geom, correct_dm) run
unmodified: every array it would upload to the GPU simulator is captured here, and the few
compute calls it makes during init (``comp_oneactu``, ``raytrace``, ``slopes_geom``) are served
by the numpy oracle (oracle/aoframe.py).
"""
import sys
import types

import numpy as np

from .. import aoframe


class Context:
    active_device = 0

    @staticmethod
    def get_instance_1gpu(dev):
        return Context()

    @staticmethod
    def get_instance_ngpu(n, devs):
        return Context()


class Telescope:
    def __init__(self, context, n, npup, pupil, n_m, mpupil):
        self.spupil = np.array(pupil)
        self.mpupil = np.array(mpupil)

    def set_phase_ab_M1(self, a):
        self.phase_ab_M1 = a

    def set_phase_ab_M1_m(self, a):
        self.phase_ab_M1_m = a


class Atmos:
    def __init__(self, context, nscreens, r0, r0_layers, dim_screens, stencil_size, alt,
                 windspeed, winddir, deltax, deltay, device):
        self.rec = dict(nscreens=int(nscreens), r0=float(r0), r0_layers=np.array(r0_layers),
                        dim_screens=np.array(dim_screens), stencil_size=np.array(stencil_size),
                        alt=np.array(alt), windspeed=np.array(windspeed),
                        winddir=np.array(winddir), deltax=np.array(deltax),
                        deltay=np.array(deltay))
        self.layers = {}

    def init_screen(self, i, A, B, istx, isty, seed):
        self.layers[int(i)] = dict(A=np.array(A), B=np.array(B), istx=np.array(istx),
                                   isty=np.array(isty), seed=int(seed))


class _Dm:
    def __init__(self, type_, alt, dim, ntotact, influsize, ninflupos, n_npts, push4imat):
        self.type = type_
        self.alt = alt
        self.dim = int(dim)
        self.ntotact = int(ntotact)
        self.influsize = int(influsize)
        self.push4imat = push4imat
        self.d_shape = np.zeros((self.dim, self.dim), np.float32)

    def pzt_loadarrays(self, influ, influpos, ninflu, influstart, i1, j1):
        self.influ = np.array(influ)
        self.influpos = np.array(influpos)
        self.ninflu = np.array(ninflu)
        self.influstart = np.array(influstart)
        self.i1 = np.array(i1)
        self.j1 = np.array(j1)

    def tt_loadarrays(self, influ):
        self.influ = np.array(influ)

    def reset_shape(self):
        self.d_shape[:] = 0

    def comp_oneactu(self, i, push):
        if self.type == "pzt":
            v = np.zeros(self.ntotact, np.float32)
            v[i] = push
            self.d_shape = aoframe.pzt_shape(v, self.influ, self.i1, self.j1, self.dim)
        else:
            v = np.zeros(2, np.float32)
            v[i] = push
            self.d_shape = aoframe.tt_shape(v, self.influ)


class Dms:
    def __init__(self):
        self.d_dms = []

    def add_dm(self, context, type_, alt, dim, ntotact, influsize, ninflupos, n_npts,
               push4imat, nord, device):
        self.d_dms.append(_Dm(type_, alt, dim, ntotact, influsize, ninflupos, n_npts, push4imat))

    def remove_dm(self, idx):
        self._removed = (idx, self.d_dms.pop(idx))

    def insert_dm(self, context, type_, alt, dim, ntotact, influsize, ninflupos, n_npts,
                  push4imat, nord, dx, dy, theta, G, device, idx):
        self.d_dms.insert(idx, _Dm(type_, alt, dim, ntotact, influsize, ninflupos, n_npts,
                                   push4imat))


class _Source:
    def __init__(self, n):
        self.n = n
        self.layers = []
        self.d_phase = np.zeros((n, n), np.float32)

    def add_layer(self, type_, idx, xoff, yoff):
        self.layers.append((str(type_), int(idx), float(xoff), float(yoff)))

    def init_strehlmeter(self):
        pass

    def raytrace(self, dms, rst=0):
        if rst:
            self.d_phase[:] = 0
        for (t, idx, xoff, yoff) in self.layers:
            if t == "atmos":
                continue
            dm = dms.d_dms[idx]
            assert xoff == yoff
            self.d_phase += aoframe.crop(dm.d_shape, self.n, xoff)


class _Wfs:
    def __init__(self, sensors, i):
        self.sensors = sensors
        self.i = i
        self.d_gs = _Source(int(sensors.rec["npup"][i]))
        self.d_slopes = None

    def load_arrays(self, phasemap, hrmap, binmap, halfxy, fluxPerSub, validsubsx, validsubsy,
                    validpuppixx, validpuppixy, ftkernel):
        self.rec = dict(phasemap=np.array(phasemap), hrmap=np.array(hrmap),
                        binmap=np.array(binmap), halfxy=np.array(halfxy),
                        fluxPerSub=np.array(fluxPerSub), validsubsx=np.array(validsubsx),
                        validsubsy=np.array(validsubsy), validpuppixx=np.array(validpuppixx),
                        validpuppixy=np.array(validpuppixy), ftkernel=np.array(ftkernel))

    def slopes_geom(self, meth):
        r = self.rec
        pd = int(self.sensors.rec["nphase"][self.i])
        origins = [(int(r["phasemap"][0, k]) // self.d_gs.n, int(r["phasemap"][0, k]) % self.d_gs.n)
                   for k in range(r["phasemap"].shape[1])]
        self.d_slopes = aoframe.slopes_geom(self.d_gs.d_phase, self.sensors.telescope.mpupil,
                                            origins, pd, r["fluxPerSub"],
                                            float(self.sensors.rec["pdiam"][self.i]))


class Sensors:
    def __init__(self, context, telescope, t_wfs, nsensors, nxsub, nvalid, nPupils, npix, nphase,
                 nrebin, nfft, ntota, npup, pdiam, nphot, nphot4imat, lgs, fakecam, maxFlux,
                 max_pix_value, device, roket):
        self.telescope = telescope
        self.rec = dict(t_wfs=list(t_wfs), nxsub=np.array(nxsub), nvalid=np.array(nvalid),
                        npix=np.array(npix), nphase=np.array(nphase), nrebin=np.array(nrebin),
                        nfft=np.array(nfft), ntota=np.array(ntota), npup=np.array(npup),
                        pdiam=np.array(pdiam), nphot=np.array(nphot),
                        nphot4imat=np.array(nphot4imat))
        self.d_wfs = [_Wfs(self, i) for i in range(int(nsensors))]

    def initgs(self, xpos, ypos, Lambda, mag, zerop, size, noise, seed, G, thetaML, dx, dy):
        self.gs = dict(xpos=np.array(xpos), ypos=np.array(ypos), Lambda=np.array(Lambda),
                       mag=np.array(mag), zerop=zerop, size=np.array(size),
                       noise=np.array(noise), seed=np.array(seed))


class Target:
    def __init__(self, ctxt, telescope, ntargets, xpos, ypos, Lambda, mag, zerop, sizes, Npts,
                 device):
        self.rec = dict(xpos=np.array(xpos), ypos=np.array(ypos), Lambda=np.array(Lambda),
                        mag=np.array(mag), zerop=zerop, sizes=np.array(sizes), Npts=float(Npts))
        self.d_targets = [_Source(int(s)) for s in sizes]


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("not provided by the recording fake")


def install():
    carma = types.ModuleType("carmaWrap")
    carma.context = Context
    sys.modules["carmaWrap"] = carma
    sutra = types.ModuleType("sutraWrap")
    for name, cls in dict(Atmos=Atmos, Telescope=Telescope, Dms=Dms, Sensors=Sensors,
                          Target=Target).items():
        setattr(sutra, name, cls)
    for name in ("Rtc_FFF", "Gamora", "Groot"):
        setattr(sutra, name, _Unavailable)
    sys.modules["sutraWrap"] = sutra
