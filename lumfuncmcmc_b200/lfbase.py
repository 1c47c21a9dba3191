"""Shared host-side machinery of the two drop-in classes (``LumFuncMCMC``, ``LumFuncMCMCz``).

Everything here runs once per fit or once per posterior summary; the per-step arithmetic is in the CUDA engine.
The set-up follows the reference's formulas exactly (same NumPy/SciPy calls on the same operands), because its
outputs -- per-source luminosities, the D_L / dV/dz interpolants, the bicubic Omega table, the quadrature grid --
ARE the engine's inputs, and drop-in parity is defined on them.
"""
import logging
import os
import time

import numpy as np
from scipy.integrate import quad
from scipy.interpolate import RectBivariateSpline

from . import VmaxLumFunc as V
from .cosmology import cosmo as _cosmo
from .setup_gpu import GPU_MIN_POINTS, LinearTable, cosmo_distances, gpu_count

MPC_CM = 3.086e24          # the reference's Mpc -> cm constant (lumfuncmcmc.py:70)


def TrueLumFunc(logL, alpha, logLstar, logphistar):
    """Schechter luminosity function per dex, Mpc^-3 dex^-1 (reference lumfuncmcmc.py:25-44)."""
    x = logL - logLstar
    return np.log(10.0) * 10 ** logphistar * 10 ** (x * (alpha + 1)) * np.exp(-10 ** x)


def Omega(logL, z, dLzfunc, Omega_0, Flim, alpha, fcmin=0.1):
    """Effective solid angle [sr] in which a source of luminosity 10**logL at redshift z is detectable: area
    times the (modified) Fleming completeness at its flux (reference lumfuncmcmc.py:47-70)."""
    L = 10 ** logL
    return Omega_0 / V.sqarcsec * V.fleming(L / (4.0 * np.pi * (MPC_CM * dLzfunc(z)) ** 2), Flim, alpha, fcmin)


class LFBase:
    """Set-up tables, engine plumbing, sampling and posterior summaries common to both models."""

    logger_name = 'lumfuncmcmc'
    engine_kind_default = 'free'

    # ------------------------------------------------------------------ set-up tables
    def _concat_inputs(self, z, flux, flux_e, lum, lum_e):
        self.z = np.concatenate(z)
        self.zmin, self.zmax = np.min(self.z), np.max(self.z)          # = min(self.z), max(self.z) without the Python-level loop
        self._flux_in, self._flux_e_in, self._lum_in, self._lum_e_in = flux, flux_e, lum, lum_e

    def defineFlimOmArr(self):
        """Per-source copies of the field's F50 and area; the area copy is integer-typed as in the reference
        (lumfuncmcmc.py:283-288), i.e. truncated toward zero."""
        n = self.field_ind[-1]
        self.Flims_arr, self.Omega_0_arr = np.zeros(n), np.zeros(n, dtype=int)
        for k in range(self.nfields):
            sl = slice(self.field_ind[k], self.field_ind[k + 1])
            self.Flims_arr[sl] = self.Flim[k]
            self.Omega_0_arr[sl] = self.Omega_0[k]

    def getFlim(self):
        for k in range(self.nfields):
            self.Flims_arr[self.field_ind[k]:self.field_ind[k + 1]] = self.Flim[k]

    def _distance_tables(self, roots):
        """D_L and dV/dz/dOmega linear interpolants on N knots over [0.95 zmin, 1.05 zmax], exact D_L per source,
        and the per-field minimum-luminosity curves (reference lumfuncmcmc.py:180-202)."""
        zint = np.linspace(0.95 * self.zmin, 1.05 * self.zmax, len(self.z))
        dev = getattr(self, 'device', 0)
        if len(self.z) >= GPU_MIN_POINTS and gpu_count() > 0:
            # the three O(N) cosmology passes on the GPU (same arithmetic as cosmology.py)
            self.DL = cosmo_distances(_cosmo, self.z, device=dev, want_dv=False)[0]
            DLarr, dVdzarr = cosmo_distances(_cosmo, zint, device=dev)
        else:
            self.DL = _cosmo.luminosity_distance(self.z)
            DLarr = _cosmo.luminosity_distance(zint)
            dVdzarr = _cosmo.differential_comoving_volume(zint)
        # interp1d(kind='linear') evaluates numpy.interp; LinearTable does the same (on the GPU for per-source arrays)
        self.DLf, self.dVdzf = LinearTable(zint, DLarr, dev), LinearTable(zint, dVdzarr, dev)
        self.minlumf = []
        for k in range(self.nfields):
            if self.min_comp_frac <= 0.001:
                minlum = np.zeros_like(DLarr)
            else:
                minlum = np.log10(4.0 * np.pi * (DLarr * MPC_CM) ** 2 * roots[k])
            self.minlumf.append(LinearTable(zint, minlum, dev))

    def _fluxes_and_luminosities(self):
        """flux <-> log-luminosity with first-order error propagation (reference lumfuncmcmc.py:165-173, 251-270)."""
        area = 4.0 * np.pi * (self.DL * MPC_CM) ** 2
        if self._flux_in is not None:
            self.flux = 1.0e-17 * np.concatenate(self._flux_in)
            if self._flux_e_in is not None:
                self.flux_e = 1.0e-17 * np.concatenate(self._flux_e_in)
        else:
            self.lum, self.lum_e = np.concatenate(self._lum_in), np.concatenate(self._lum_e_in)
            if self.lum_e is not None:
                L = 10 ** self.lum
                self.flux, self.flux_e = L / area, np.abs(L * np.log(10)) * self.lum_e / np.abs(area)
            else:
                self.flux, self.flux_e = 10 ** self.lum / area, None
        if self._lum_in is None:
            if self.flux_e is not None:          # AttributeError if flux came without errors, as in the reference
                nominal = area * self.flux
                self.lum = np.log10(nominal)
                self.lum_e = np.abs((np.abs(area) * self.flux_e) / (nominal * np.log(10.0)))
            else:
                self.lum, self.lum_e = np.log10(area * self.flux), None

    def setOmegaLz(self, size=501):
        """Bicubic table of Omega(logL, z) per field on a size x size grid (reference lumfuncmcmc.py:204-215)."""
        logL = np.linspace(self.Lc, self.Lh, size)
        zarr = np.linspace(0.95 * self.zmin, 1.05 * self.zmax, size)
        self.Omegaf = []
        table = np.empty((size, size))
        for k in range(self.nfields):
            # row by row with a SCALAR luminosity: NumPy's scalar and array power loops may round differently, and
            # the reference tabulates with scalars (lumfuncmcmc.py:212-214)
            for i in range(size):
                table[i] = Omega(logL[i], zarr, self.DLf, self.Omega_0[k], 1.0e-17 * self.Flim[k], self.alpha,
                                 self.fcmin)
            self.Omegaf.append(RectBivariateSpline(logL, zarr, table))

    def _quadrature_grid(self, size_ln):
        """(logL, z) trapezoid grid per field, the tabulated-Omega integrand factor and the per-source Omega
        (reference lumfuncmcmc.py:217-235).  The reference appends ONE array object for every field while it keeps
        overwriting it, so after set-up every ``logL[k]`` is the LAST field's grid, while ``integ_part[k]`` was
        tabulated on field k's own grid; identical grids when min_comp_frac <= 0.001.  That
        aliasing is part of the reference's numbers and is kept."""
        S = self.size_ln = size_ln
        self.zarr = np.linspace(self.zmin, self.zmax, S)
        self.DL_zarr = self.DLf(self.zarr)
        self.volume_part = self.dVdzf(self.zarr)
        self.zarr_rep = np.repeat(self.zarr[None], S, axis=0)
        self.integ_part = []
        lum_floor = np.min(self.lum)
        for k in range(self.nfields):
            lo = self.minlumf[k](self.zarr)
            lo[lo < lum_floor] = lum_floor
            self.logLi = np.empty((S, S))
            for i in range(S):
                self.logLi[:, i] = np.linspace(lo[i], self.Lh, S)
            self.integ_part.append(self.volume_part * self.Omegaf[k].ev(self.logLi, self.zarr_rep))
        self.logL = [self.logLi] * self.nfields
        if len(self.lum) >= GPU_MIN_POINTS and gpu_count() > 0:
            # the O(N) pass of the set-up chain on the GPU: same formula and order of operations, libdevice transcendentals
            # (within ~1e-15 of NumPy's; the engine input either way)
            from .setup_gpu import omega_sources
            om0_field = [int(self.Omega_0_arr[self.field_ind[k]]) if self.field_ind[k + 1] > self.field_ind[k] else 0
                         for k in range(self.nfields)]
            self.Om_arr = omega_sources(self.lum, self.z, self.field_ind, om0_field, self.Flim, self.alpha, self.fcmin, self.DLf,
                                        device=getattr(self, 'device', 0))
        else:
            self.Om_arr = Omega(self.lum, self.z, self.DLf, self.Omega_0_arr, 1.0e-17 * self.Flims_arr, self.alpha,
                                self.fcmin)

    def setup_logging(self):
        self.log = logging.getLogger(self.logger_name)
        if not len(self.log.handlers):
            handler = logging.StreamHandler()
            handler.setFormatter(logging.Formatter('[%(levelname)s - %(asctime)s] %(message)s'))
            handler.setLevel(logging.INFO)
            self.log.setLevel(logging.DEBUG)
            self.log.addHandler(handler)

    # ------------------------------------------------------------------ engine plumbing
    def engine_inputs(self):
        """The arrays the likelihood reads, as one dict (same keys as the oracle / ``LikelihoodEngine``)."""
        inp = dict(lum=self.lum, z=self.z, zint=self.DLf.x, DLarr=self.DLf.y, dVdzarr=self.dVdzf.y,
                   field_ind=np.asarray(self.field_ind, dtype=np.int64),
                   Omega_0=np.asarray(self.Omega_0, dtype=np.float64), Flim=np.asarray(self._Flim0, dtype=np.float64),
                   alpha=float(self._alpha0), fcmin=self.fcmin, logL=np.stack(self.logL), zarr=self.zarr,
                   DL_zarr=self.DL_zarr, volume_part=self.volume_part, Om_arr=self.Om_arr,
                   integ_part=np.stack(self.integ_part), Lstar_lims=self.Lstar_lims, phistar_lims=self.phistar_lims,
                   sch_al_lims=self.sch_al_lims, sch_al=float(self._sch_al0), fix_sch_al=bool(self.fix_sch_al))
        return inp

    def _engine(self, kind):
        eng = self._engines.get(kind)
        if eng is None:
            from .engine import LikelihoodEngine
            # opt-in: model.compress = True (or LF_COMPRESS=1) evaluates the source sum on weighted pseudo-sources
            compress = (bool(getattr(self, 'compress', False)) or os.environ.get('LF_COMPRESS', '') == '1') and kind != 'fixed'
            eng = LikelihoodEngine(self.engine_inputs(), kind, device=self.device, compress=compress)
            self._engines[kind] = eng
        return eng

    def close(self):
        for eng in self._engines.values():
            eng.close()
        self._engines = {}

    # ------------------------------------------------------------------ sampling
    def _device_sampler_engine(self, func):
        """Engine whose parameter rows ARE the sampled vector for ``func`` (None: host sampler only)."""
        return None

    def _run_sampler(self, func):
        """emcee-style ensemble run handing the GPU a whole half-ensemble per call (reference lumfuncmcmc.py:479-513)."""
        from .sampler import DeviceEnsembleSampler, EnsembleSampler
        self.log.info('Fitting Schechter model to true luminosity function using emcee')
        pos = self.get_init_walker_values()
        ndim = pos.shape[1]
        start = time.time()
        # sampler_backend = 'device' (attribute or LF_SAMPLER=device): the whole run stays on the GPU whenever the sampled
        # parameter vector is the engine's own row layout; otherwise (and by default) the host sampler drives the engine
        backend = getattr(self, 'sampler_backend', None) or os.environ.get('LF_SAMPLER', 'host')
        eng = self._device_sampler_engine(func) if backend == 'device' else None
        if eng is not None and eng.ndim == ndim:
            sampler = DeviceEnsembleSampler(self.nwalkers, ndim, eng)
            sampler.run_mcmc(pos, self.nsteps, rstate0=np.random.get_state())
            self.set_parameters_from_list(sampler.chain[-1, -1, :])
        else:
            sampler = EnsembleSampler(self.nwalkers, ndim, func, vectorize=True)
            sampler.run_mcmc(pos, self.nsteps, rstate0=np.random.get_state())
        elapsed = time.time() - start
        self.log.info("Total time taken: %0.2f s" % elapsed)
        self.log.info("Time taken per step per walker: %0.2f ms" % (elapsed / (self.nsteps) * 1000. / self.nwalkers))
        tau = np.max(sampler.acor)
        burnin_step = int(tau * 3)
        if burnin_step > self.nsteps // 2:
            burnin_step = self.nsteps // 2
        self.log.info("Mean acceptance fraction: %0.2f" % (np.mean(sampler.acceptance_fraction)))
        self.log.info("AutoCorrelation Steps: %i, Number of Burn-in Steps: %i" % (np.round(tau), burnin_step))
        new_chain = np.zeros((self.nwalkers, self.nsteps, ndim + 1))
        new_chain[:, :, :-1] = sampler.chain
        self.chain = sampler.chain
        new_chain[:, :, -1] = sampler.lnprobability
        self.samples = new_chain[:, burnin_step:, :].reshape((-1, ndim + 1))
        self.log.info("Shape of self.samples")
        self.log.info(self.samples.shape)
        self.log.info("Median lnprob: %.5f; Max lnprob: %.5f" % (np.median(sampler.lnprobability),
                                                               np.amax(sampler.lnprobability)))
        self.sampler = sampler

    # ------------------------------------------------------------------ posterior summaries
    def _lnprob_selection(self, lnprobcut, drop_lnprob):
        """Samples within ``lnprobcut`` of the maximum; the cut doubles until a quarter of the samples survive
        (reference lumfuncmcmc.py:548-553, 620-625, 655-660)."""
        nsamples = []
        while len(nsamples) < len(self.samples) // 4:
            keep = self.samples[:, -1] > (np.max(self.samples[:, -1], axis=0) - lnprobcut)
            nsamples = self.samples[keep, :-1] if drop_lnprob else self.samples[keep, :]
            lnprobcut *= 2.0
        return nsamples

    def add_fitinfo_to_table(self, percentiles, start_value=1, lnprobcut=7.5):
        """Percentiles of every parameter into the last table row (reference lumfuncmcmc.py:653-667)."""
        nsamples = self._lnprob_selection(lnprobcut, drop_lnprob=True)
        self.log.info("Number of table entries: %d" % (len(self.table[0])))
        n = len(percentiles)
        for i, per in enumerate(percentiles):
            for j, v in enumerate(np.percentile(nsamples, per, axis=0)):
                self.table[-1][(i + start_value + j * n)] = v

    # ------------------------------------------------------------------ 1/V_eff luminosity function
    _phifunc, _phi_engine, _phi_gen = None, None, -1

    @property
    def phifunc(self):
        """Per-source 1/V_eff weights (reference attribute of the same name).  ``VeffLF`` leaves them on the GPU; they are
        copied to the host the first time this attribute is read."""
        if self._phifunc is None and self._phi_engine is not None:
            if self._phi_engine._phi_gen != self._phi_gen:
                raise RuntimeError("the 1/V_eff weights are no longer resident on the device (another sample was binned "
                                   "on this engine); call VeffLF() again")
            self._phifunc = self._phi_engine.veff_phi()
        return self._phifunc

    @phifunc.setter
    def phifunc(self, value):
        self._phifunc, self._phi_engine = value, None

    def _veff_volumes_host(self, root_per_field):
        """The reference's per-source loop (lumfuncmcmc.py:521-524): one fsolve (``V.getMaxz``) and one QUADPACK integral
        per source.  Kept as the checked oracle of the device path (tests); O(N) Python, hours at 1e7 sources."""
        n = len(self.flux)
        root_per_source = np.repeat(np.asarray(root_per_field, dtype=np.float64), np.diff(np.asarray(self.field_ind)))
        vol, valid, zmaxs = np.ones(n), np.zeros(n, dtype=np.uint8), np.zeros(n)
        for i in range(n):
            zmaxval = min(self.zmax, V.getMaxz(10 ** self.lum[i], root_per_source[i]))
            zmaxs[i] = zmaxval
            if zmaxval > self.zmin:
                vol[i] = quad(self.dVdzf, self.zmin, zmaxval)[0]
                valid[i] = 1
        return zmaxs, vol, valid

    def _veff(self, root_per_field):
        """Per-source 1/V_eff weights, binned LF and bootstrap errors on the GPU (reference lumfuncmcmc.py:515-525).

        The catalogue (flux, lum) is uploaded once per engine and stays resident: VeffLF runs after every fit and again
        for each posterior summary with new completeness parameters.  With min_comp_frac <= 0.001 every source integrates
        dV/dz over the whole [zmin, zmax]: one shared QUADPACK integral of the interpolant, exactly the number the
        reference computes N times.  Otherwise each source's upper limit is where its luminosity drops to the field's
        minimum flux and its volume is the integral of dVdzf to that limit: the reference's per-source fsolve + QUADPACK
        loop becomes one kernel (``lf_veff_volumes``: Newton on D_L, exact integral of the linear interpolant), equal to
        the loop within its own solver tolerances (1.5e-8; :meth:`_veff_volumes_host` is the loop itself)."""
        sum_Omega = sum(self.Omega_0)
        n = len(self.flux)
        eng = self._veff_engine()
        if getattr(self, '_veff_sample_on', None) is not eng or eng._sample_owner is not self:
            eng.veff_set_sample(self.flux, self.lum, self.field_ind)
            eng._sample_owner, self._veff_sample_on, self._veff_table_on = self, eng, None
        lum_lo, lum_hi = np.min(self.lum), np.max(self.lum)
        edges = np.linspace(lum_lo * 1.001, lum_hi, self.nbins + 1)
        if self.min_comp_frac <= 0.001 and self.zmax > self.zmin:
            vol_int = quad(self.dVdzf, self.zmin, self.zmax)[0]
            _, counts, sums = eng.veff_bin_resident(self.Flim, self.alpha, self.fcmin, sum_Omega, vol_int, edges)
        elif self.min_comp_frac <= 0.001:
            # degenerate catalogue (one redshift): no source has a volume, every weight is 0 (lumfuncmcmc.py:522-523)
            _, counts, sums = eng.veff_bin(self.flux, self.lum, self.field_ind, self.Flim, self.alpha, self.fcmin, sum_Omega,
                                           1.0, edges, valid=np.zeros(n, dtype=np.uint8), want_phi=False)
            self._veff_sample_on = None
        else:
            if self._veff_table_on is not eng:
                eng.veff_set_volume_table(_cosmo, self.dVdzf.x, self.dVdzf.y)
                self._veff_table_on = eng
            fmin = np.asarray(root_per_field, dtype=np.float64)          # minimum flux per field (cgs)
            if fmin.shape != (self.nfields,):
                raise ValueError("one minimum flux per field expected")
            eng.veff_volumes(self.zmin, self.zmax, float(_cosmo.luminosity_distance(self.zmin)),
                             float(_cosmo.luminosity_distance(self.zmax)), fmin)
            _, counts, sums = eng.veff_bin_resident(self.Flim, self.alpha, self.fcmin, sum_Omega, 1.0, edges,
                                                    device_volumes=True)
        self._phifunc, self._phi_engine, self._phi_gen = None, eng, eng._phi_gen
        self.Lavg, self.lfbinorig, self.var, self.bincounts = V.getBootErrLog(
            self.lum, None, self.zmin, self.zmax, self.nboot, self.nbins, Fmin=1.0e-17 * np.max(self.Flim),
            engine=eng, return_counts=True, rng=getattr(self, 'boot_rng', None) or os.environ.get('LF_BOOT_RNG', 'auto'),
            Lrange=(lum_lo, lum_hi))

    def _veff_engine(self):
        if self._engines:
            return next(iter(self._engines.values()))
        if getattr(self, '_veff_only', None) is None:
            from .engine import VeffEngine
            self._veff_only = VeffEngine(device=self.device)
        return self._veff_only
