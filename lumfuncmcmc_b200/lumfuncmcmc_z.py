"""Drop-in ``LumFuncMCMCz``: Schechter function whose log L* and log phi* are quadratics in redshift through three
pivot redshifts (reference lumfuncmcmc_z.py:118-609), likelihood on the B200 engine.  Completeness is always fixed
(tabulated Omega); theta = [L1, L2, L3, phi1, phi2, phi3, (alpha_s)].
"""
import numpy as np
from scipy.optimize import fsolve

from . import VmaxLumFunc as V
from .lfbase import LFBase, Omega, TrueLumFunc  # noqa: F401


def getQuadCoef(y1, y2, y3, z1, z2, z3):
    """Coefficients (a, b, c) of y = a z^2 + b z + c through (z1, y1), (z2, y2), (z3, y3)
    (reference lumfuncmcmc_z.py:26-43; same operation order, so the same doubles)."""
    a = ((y3 - y1) + (y2 - y1) * (z1 - z3) / (z2 - z1)) / (z3 ** 2 - z1 ** 2 + (z2 ** 2 - z1 ** 2) * (z1 - z3) / (z2 - z1))
    b = (y2 - y1 - a * (z2 ** 2 - z1 ** 2)) / (z2 - z1)
    c = y1 - a * z1 ** 2 - b * z1
    return a, b, c


def schechter_z(L, z, al, L1, L2, L3, phi1, phi2, phi3, z1, z2, z3):
    """Schechter function at (log L, z) with L*(z), phi*(z) interpolated quadratically between the pivots
    (reference lumfuncmcmc_z.py:45-67)."""
    aphi, bphi, cphi = getQuadCoef(phi1, phi2, phi3, z1, z2, z3)
    alum, blum, clum = getQuadCoef(L1, L2, L3, z1, z2, z3)
    return TrueLumFunc(L, al, alum * z ** 2 + blum * z + clum, aphi * z ** 2 + bphi * z + cphi)


class LumFuncMCMCz(LFBase):
    logger_name = 'lumfuncmcmc_z'

    def __init__(self, z, flux=None, flux_e=None, Flim=[2.35, 3.12, 2.20, 2.86, 2.85], alpha=3.5, line_name="OIII",
                 line_plot_name=r'[OIII] $\lambda 5007$', lum=None, lum_e=None,
                 Omega_0=[100.0, 100.0, 100.0, 100.0, 100.0], nbins=50, nboot=100, sch_al=-1.6,
                 sch_al_lims=[-3.0, 1.0], Lstar=42.5, Lstar_lims=[41.0, 45.0], phistar=-3.0, phistar_lims=[-8.0, 5.0],
                 Lc=40.0, Lh=46.0, nwalkers=100, nsteps=1000, fcmin=0.1, min_comp_frac=0.5, field_names=None,
                 field_ind=None, z1=1.20, z2=1.53, z3=1.86, fix_sch_al=False, device=0):
        """Same arguments as the reference (lumfuncmcmc_z.py:119-188); ``z1, z2, z3`` are the pivot redshifts."""
        self._concat_inputs(z, flux, flux_e, lum, lum_e)
        self.z1, self.z2, self.z3 = z1, z2, z3
        self.fcmin, self.min_comp_frac = fcmin, min_comp_frac
        self.Flim = Flim
        self.fields, self.nfields = field_names, len(self.Flim)
        self.field_ind = field_ind
        self.alpha = alpha
        self.line_name, self.line_plot_name = line_name, line_plot_name
        self.Lc, self.Lh = Lc, Lh
        self.Omega_0 = Omega_0
        self.fix_sch_al = fix_sch_al
        self.nbins, self.nboot = nbins, nboot
        self.sch_al, self.sch_al_lims = sch_al, sch_al_lims
        self.Lstar, self.Lstar_lims = Lstar, Lstar_lims
        self.phistar, self.phistar_lims = phistar, phistar_lims
        # start values of the pivots: drawn from the global stream, as the reference does (lumfuncmcmc_z.py:206-207)
        self.L1, self.L2, self.L3 = np.random.uniform(self.Lstar_lims[0] + 0.5, self.Lstar_lims[-1] - 0.5, 3)
        self.phi1, self.phi2, self.phi3 = np.random.uniform(self.phistar_lims[0] + 3, self.phistar_lims[-1] - 3, 3)
        self.nwalkers, self.nsteps = nwalkers, nsteps
        self.device = device
        self._engines = {}
        self._Flim0, self._alpha0, self._sch_al0 = list(Flim), alpha, sch_al
        self.getRoot()
        self.defineFlimOmArr()
        self.setDLdVdz()
        self._fluxes_and_luminosities()
        self.setOmegaLz()
        self.allind = np.arange(len(self.lum))
        self.setlnsimple()
        self.setup_logging()

    # ------------------------------------------------------------------ set-up
    def getRoot(self):
        """Per-field flux at which completeness = min_comp_frac (reference lumfuncmcmc_z.py:292-297)."""
        self.roots_ln = np.array([])
        for i in range(self.nfields):
            root = fsolve(lambda x: V.fleming(x, 1.0e-17 * self.Flim[i], self.alpha, self.fcmin) - self.min_comp_frac,
                          [1.0e-17 * self.Flim[i]])[0]
            self.roots_ln = np.append(self.roots_ln, root)

    def defineFlimOmArr(self):
        super().defineFlimOmArr()
        self.roots_arr = np.zeros(self.field_ind[-1])
        for k in range(self.nfields):
            self.roots_arr[self.field_ind[k]:self.field_ind[k + 1]] = self.roots_ln[k]

    def setDLdVdz(self):
        self._distance_tables(self.roots_ln)

    def setlnsimple(self):
        self._quadrature_grid(201)

    # ------------------------------------------------------------------ parameters and prior
    def set_parameters_from_list(self, input_list):
        self.L1, self.L2, self.L3 = input_list[0], input_list[1], input_list[2]
        self.phi1, self.phi2, self.phi3 = input_list[3], input_list[4], input_list[5]
        if not self.fix_sch_al:
            self.sch_al = input_list[6]

    def lnprior(self):
        """Flat prior: alpha_s inclusive in its box (when sampled), the six pivots strictly inside theirs
        (reference lumfuncmcmc_z.py:343-362)."""
        inside = True
        if not self.fix_sch_al:
            inside = bool((self.sch_al >= self.sch_al_lims[0]) * (self.sch_al <= self.sch_al_lims[1]))
        for v in (self.L1, self.L2, self.L3):
            inside = inside and bool((v > self.Lstar_lims[0]) * (v < self.Lstar_lims[1]))
        for v in (self.phi1, self.phi2, self.phi3):
            inside = inside and bool((v > self.phistar_lims[0]) * (v < self.phistar_lims[1]))
        return 0.0 if inside else -np.inf

    def _current_theta(self):
        vals = [self.L1, self.L2, self.L3, self.phi1, self.phi2, self.phi3]
        if not self.fix_sch_al:
            vals.append(self.sch_al)
        return np.array(vals, dtype=np.float64)[None, :]

    # ------------------------------------------------------------------ likelihood
    def engine_inputs(self):
        inp = super().engine_inputs()
        inp.update(z1=self.z1, z2=self.z2, z3=self.z3, sch_al=float(self.sch_al))
        return inp

    def _z_engine(self):
        # with a fixed alpha_s the engine bakes its value in: rebuild if the attribute was changed since
        eng = self._engines.get('z')
        if eng is not None and self.fix_sch_al and self._baked_sch_al != float(self.sch_al):
            eng.close()
            del self._engines['z']
        self._baked_sch_al = float(self.sch_al)
        return self._engine('z')

    def lnlike(self):
        """ln L at the current attributes (reference lumfuncmcmc_z.py:364-376).  No prior."""
        return float(self._z_engine().lnlike(self._current_theta())[0])

    def lnprob(self, theta):
        """ln prior + ln likelihood; (ndim,) -> float, (W, ndim) -> (W,) (reference lumfuncmcmc_z.py:378-392)."""
        th = np.asarray(theta, dtype=np.float64)
        out = self._z_engine().lnprob(np.ascontiguousarray(np.atleast_2d(th)))
        if th.ndim == 1:
            self.set_parameters_from_list(theta)
            return float(out[0])
        self.set_parameters_from_list(th[-1])
        return out

    # ------------------------------------------------------------------ sampler host
    def get_init_walker_values(self, num=None):
        lims = [self.Lstar_lims] * 3 + [self.phistar_lims] * 3
        if not self.fix_sch_al:
            lims.append(self.sch_al_lims)
        lims = np.array(lims, dtype=np.float64)
        num = self.nwalkers if num is None else num
        return np.random.rand(num, len(lims)) * (lims[:, 1] - lims[:, 0]) + lims[:, 0]

    def get_param_names(self):
        names = [r'$\log {\rm{L}}1_*$', r'$\log {\rm{L}}2_*$', r'$\log {\rm{L}}3_*$',
                 r'$\log \phi1_*$', r'$\log \phi2_*$', r'$\log \phi3_*$']
        if not self.fix_sch_al:
            names.append(r'$\alpha$')
        return names

    def get_params(self):
        vals = list(self._current_theta()[0])
        self.nfreeparams = len(vals)
        return vals

    def _device_sampler_engine(self, func):
        return self._z_engine()

    def fit_model(self):
        self._run_sampler(self.lnprob)

    # ------------------------------------------------------------------ 1/V_eff and posterior summaries
    def VeffLF(self):
        """1/V_eff binned LF with bootstrap errors (reference lumfuncmcmc_z.py:470-478)."""
        self._veff(self.roots_ln)

    def _median_matrix(self, nsamples, lo_pad, hi_pad, zlen, Llen):
        """Model LF at the posterior-median parameters on a (redshift, luminosity) mesh -> ``Lout``, ``zout``,
        ``medianLF[zlen, Llen]``, then the 1/V_eff LF (reference lumfuncmcmc_z.py:507-515, 524-533)."""
        self.Lout = np.linspace(min(self.lum) - lo_pad, max(self.lum) + hi_pad, Llen)
        self.zout = np.linspace(self.zmin, self.zmax, zlen)
        self.medianLF = np.zeros((zlen, Llen))
        self.set_parameters_from_list(np.percentile(nsamples, 50.0, axis=0))
        for i in np.arange(zlen):
            self.medianLF[i] = schechter_z(self.Lout, self.zout[i], self.sch_al, self.L1, self.L2, self.L3, self.phi1,
                                           self.phi2, self.phi3, self.z1, self.z2, self.z3)
        self.VeffLF()

    def set_median_fit(self, lnprobcut=7.5, zlen=100, Llen=100):
        """Posterior-median model on a mesh and the 1/V_eff LF, without a figure (reference lumfuncmcmc_z.py:480-515)."""
        nsamples = self._lnprob_selection(lnprobcut, drop_lnprob=True)
        self.log.info("Shape of nsamples (with a lnprobcut applied)")
        self.log.info(nsamples.shape)
        self._median_matrix(nsamples, 0.2, 0.2, zlen, Llen)

    def triangle_plot(self, outname, lnprobcut=7.5, imgtype='png'):
        """Corner plot with the LF-vs-redshift panel (reference lumfuncmcmc_z.py:524-593).  The data products
        (``Lout``, ``zout``, ``medianLF``, ``VeffLF``) are always computed; the figure needs matplotlib + corner."""
        nsamples = self._lnprob_selection(lnprobcut, drop_lnprob=False)
        self.log.info("Shape of nsamples (with a lnprobcut applied)")
        self.log.info(nsamples.shape)
        self._median_matrix(nsamples, 0.08, 0.01, 100, 100)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import corner
        except ImportError:
            self.log.info("matplotlib/corner not available: skipping the figure")
            return
        names = self.get_param_names()
        nd = len(nsamples[0])
        fs = 11 + int(round(0.75 * nd))
        fig = corner.corner(nsamples[:, :-1], labels=names, range=[.95] * len(names), label_kwargs={"fontsize": fs},
                            show_titles=True, title_kwargs={"fontsize": fs - 2}, quantiles=[0.16, 0.5, 0.84], bins=30)
        w = fig.get_figwidth()
        if nd >= 4:
            fig.set_figwidth(w - (nd - 13) * 0.025 * w)
            box = [0.44 - 0.008 * (nd - 4), 0.78 - 0.001 * (nd - 4), 0.48 + 0.008 * (nd - 4), 0.19 + 0.001 * (nd - 4)]
        else:
            box = [0.67, 0.75, 0.32, 0.23]
        ax = fig.add_subplot(3, 1, 1)
        ax.set_position(box)
        ax.set_yscale('log')
        ax.set_xlabel(r"$\log$ L (erg s$^{-1}$)")
        ax.set_ylabel(r"$\phi_{\rm{true}}$ (Mpc$^{-3}$ dex$^{-1}$)")
        LL, zz = np.meshgrid(self.Lout, self.zout)
        im = ax.pcolormesh(LL, self.medianLF, zz, shading='auto', cmap='viridis')
        xmax = min(max(self.L1, self.L2, self.L3) + 0.5, self.Lout.max())
        ax.set_xlim(right=xmax)
        fig.colorbar(im, ax=ax, label='Redshift')
        fig.savefig("%s.%s" % (outname, imgtype), dpi=200)
        plt.close(fig)
