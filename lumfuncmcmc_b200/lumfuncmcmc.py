"""Drop-in ``LumFuncMCMC``: the reference's class surface (constructor keywords, method and attribute names,
reference lumfuncmcmc.py:72-667) with the likelihood evaluated by the B200 engine.

What changes for a user of the reference:
  * ``lnprob`` / ``lnprob_fix_comp`` also accept a 2-D ``(W, ndim)`` array and return ``(W,)`` -- one GPU call for
    the whole ensemble (a 1-D theta is a batch of one and still returns a float and updates ``self``);
  * ``fit_model`` drives a vectorised ensemble sampler (``lumfuncmcmc_b200.sampler``) instead of one Python call
    per walker;
  * ``VeffLF`` computes weights, binning and bootstrap replicates on the GPU;
  * one extra keyword, ``device`` (CUDA ordinal).
There is no CPU path: evaluating the likelihood without the built CUDA library or without a GPU raises.
"""
import numpy as np
from scipy.interpolate import RectBivariateSpline
from scipy.optimize import fsolve

from . import VmaxLumFunc as V
from .lfbase import LFBase, Omega, TrueLumFunc  # noqa: F401  (module-level names of the reference)


class LumFuncMCMC(LFBase):
    logger_name = 'lumfuncmcmc'

    def __init__(self, z, flux=None, flux_e=None, Flim=[2.35, 3.12, 2.20, 2.86, 2.85], Flim_lims=[1.0, 6.0],
                 alpha=3.5, alpha_lims=[1.0, 6.0], line_name="OIII", line_plot_name=r'[OIII] $\lambda 5007$',
                 lum=None, lum_e=None, Omega_0=[100.0, 100.0, 100.0, 100.0, 100.0], nbins=50, nboot=100,
                 sch_al=-1.6, sch_al_lims=[-3.0, 1.0], Lstar=42.5, Lstar_lims=[40.0, 45.0], phistar=-3.0,
                 phistar_lims=[-8.0, 5.0], Lc=40.0, Lh=46.0, nwalkers=100, nsteps=1000, fix_sch_al=False,
                 fcmin=0.1, fix_comp=False, min_comp_frac=0.5, field_names=None, field_ind=None, diff_rand=True,
                 device=0):
        """Same arguments as the reference (lumfuncmcmc.py:73-141): per-field lists ``z`` and ``flux``/``flux_e``
        (1e-17 erg/cm^2/s) or ``lum``/``lum_e``; per-field ``Flim`` (50% completeness flux) and ``Omega_0``
        (arcsec^2); completeness slope ``alpha``; Schechter start values and prior boxes; integration limits
        ``Lc``, ``Lh``; sampler size; ``fix_sch_al`` / ``fix_comp`` to remove parameters from the fit;
        ``min_comp_frac``; ``field_ind`` = cumulative per-field offsets into the concatenated arrays."""
        self._concat_inputs(z, flux, flux_e, lum, lum_e)
        self.fcmin, self.min_comp_frac = fcmin, min_comp_frac
        self.Flim, self.Flim_lims = Flim, Flim_lims
        self.fields, self.nfields = field_names, len(self.Flim)
        self.field_ind = field_ind
        self.alpha, self.alpha_lims = alpha, alpha_lims
        self.line_name, self.line_plot_name = line_name, line_plot_name
        self.Lc, self.Lh = Lc, Lh
        self.Omega_0 = Omega_0
        self.nbins, self.nboot = nbins, nboot
        self.sch_al, self.sch_al_lims = sch_al, sch_al_lims
        self.Lstar, self.Lstar_lims = Lstar, Lstar_lims
        self.phistar, self.phistar_lims = phistar, phistar_lims
        self.nwalkers, self.nsteps = nwalkers, nsteps
        self.fix_sch_al, self.fix_comp = fix_sch_al, fix_comp
        self.all_param_names = ['Lstar', 'phistar', 'sch_al', 'Flim', 'alpha']
        self.diff_rand = diff_rand
        self.device = device
        self._engines = {}
        self._Flim0, self._alpha0, self._sch_al0 = list(Flim), alpha, sch_al
        self.defineFlimOmArr()
        self.getRoot()
        self.setDLdVdz()
        self._fluxes_and_luminosities()
        self.setOmegaLz()
        self.roots_ln = self.rootsf.ev(self.Flim, self.alpha)
        self.allind = np.arange(len(self.lum))
        self.setlnsimple()
        self.setup_logging()

    # ------------------------------------------------------------------ set-up (reference method names)
    def getRoot(self, size=201):
        """Spline of the flux at which completeness = min_comp_frac over the (F50, alpha) prior box; zeros when no
        minimum completeness is imposed (reference lumfuncmcmc.py:272-281)."""
        flims = np.linspace(self.Flim_lims[0], self.Flim_lims[1], size)
        alphas = np.linspace(self.alpha_lims[0], self.alpha_lims[1], size)
        roots = np.zeros((size, size))
        if self.min_comp_frac > 0.001:
            for i in range(size):
                for j in range(size):
                    roots[i, j] = fsolve(lambda x: V.fleming(x, 1.0e-17 * flims[i], alphas[j], self.fcmin)
                                         - self.min_comp_frac, [3.0e-17])[0]
        self.rootsf = RectBivariateSpline(flims, alphas, roots)

    def setDLdVdz(self):
        self._distance_tables(self.rootsf.ev(self.Flim, self.alpha))

    def setlnsimple(self):
        self._quadrature_grid(201 if self.fix_comp else 101)

    # ------------------------------------------------------------------ parameters and prior
    def set_parameters_from_list(self, input_list):
        """theta -> attributes: [L*, phi*, (alpha_s), (F50_0..F50_{K-1}, alpha_c)] (reference lumfuncmcmc.py:320-337)."""
        K = self.nfields
        self.Lstar, self.phistar = input_list[0], input_list[1]
        nxt = 2
        if not self.fix_sch_al:
            self.sch_al = input_list[2]
            nxt = 3
        if not self.fix_comp:
            self.Flim, self.alpha = input_list[nxt:nxt + K], input_list[nxt + K]

    def lnprior(self):
        """Flat box prior on all five parameter groups, bounds inclusive (reference lumfuncmcmc.py:339-358)."""
        inside = True
        for name in self.all_param_names:
            lo, hi = getattr(self, name + '_lims')
            vals = getattr(self, name)
            for v in (vals if name == 'Flim' else [vals]):
                inside = inside and bool((v >= lo) * (v <= hi))
        return 0.0 if inside else -np.inf

    def _theta_matrix(self, theta2d, free):
        """Engine parameter rows [L*, phi*, alpha_s, (F50..., alpha_c)] from user rows, filling the parameters that
        are not being sampled from the current attributes."""
        W, K = theta2d.shape[0], self.nfields
        cols = [theta2d[:, 0], theta2d[:, 1]]
        nxt = 2
        if self.fix_sch_al:
            cols.append(np.full(W, float(self.sch_al)))
        else:
            cols.append(theta2d[:, 2])
            nxt = 3
        if free:
            if self.fix_comp:
                cols += [np.full(W, float(f)) for f in self.Flim] + [np.full(W, float(self.alpha))]
            else:
                cols += [theta2d[:, nxt + k] for k in range(K)] + [theta2d[:, nxt + K]]
        return np.ascontiguousarray(np.column_stack(cols))

    def _current_theta(self):
        vals = [self.Lstar, self.phistar]
        if not self.fix_sch_al:
            vals.append(self.sch_al)
        if not self.fix_comp:
            vals += list(self.Flim) + [self.alpha]
        return np.array(vals, dtype=np.float64)[None, :]

    def _completeness_in_prior(self):
        lo, hi = self.Flim_lims
        ok = all((f >= lo) and (f <= hi) for f in self.Flim)
        return ok and (self.alpha >= self.alpha_lims[0]) and (self.alpha <= self.alpha_lims[1])

    # ------------------------------------------------------------------ likelihood
    def engine_inputs(self):
        inp = super().engine_inputs()
        inp.update(Flim_lims=self.Flim_lims, alpha_lims=self.alpha_lims, fix_sch_al=False)
        return inp

    def lnlike(self):
        """ln L at the current attributes, completeness parameters live: sum over sources of ln(Phi*Omega) minus
        the (logL, z) integral per field (reference lumfuncmcmc.py:360-378).  No prior."""
        self.getFlim()
        return float(self._engine('free').lnlike(self._theta_matrix(self._current_theta(), True))[0])

    def lnlike_fix_comp(self):
        """Same with the tabulated Omega of the configured completeness (reference lumfuncmcmc.py:380-393)."""
        self.getFlim()
        return float(self._engine('fixed').lnlike(self._theta_matrix(self._current_theta(), False))[0])

    def _lnprob(self, theta, free):
        th = np.asarray(theta, dtype=np.float64)
        scalar = th.ndim == 1
        th2 = np.atleast_2d(th)
        out = self._engine('free' if free else 'fixed').lnprob(self._theta_matrix(th2, free))
        if not free and not self._completeness_in_prior():
            out = np.full_like(out, -np.inf)     # the reference's prior also range-checks the fixed F50 / alpha_c
        if scalar:
            self.set_parameters_from_list(theta)
            return float(out[0])
        self.set_parameters_from_list(th2[-1])
        return out

    def lnprob(self, theta):
        """ln prior + ln likelihood; ``theta`` of shape (ndim,) -> float, (W, ndim) -> (W,) (reference
        lumfuncmcmc.py:395-409).  -inf outside the prior box or when any source term underflows."""
        return self._lnprob(theta, True)

    def lnprob_fix_comp(self, theta):
        """Fixed-completeness variant (reference lumfuncmcmc.py:411-424)."""
        return self._lnprob(theta, False)

    # ------------------------------------------------------------------ sampler host
    def get_init_walker_values(self, num=None):
        """Uniform draws inside the prior box from NumPy's global stream (reference lumfuncmcmc.py:426-446)."""
        lims = [self.Lstar_lims, self.phistar_lims]
        if not self.fix_sch_al:
            lims.append(self.sch_al_lims)
        if not self.fix_comp:
            lims += [self.Flim_lims] * self.nfields + [self.alpha_lims]
        lims = np.array(lims, dtype=np.float64)
        num = self.nwalkers if num is None else num
        u = np.random.rand(num, len(lims)) if self.diff_rand else np.random.rand(num)[:, np.newaxis]
        return u * (lims[:, 1] - lims[:, 0]) + lims[:, 0]

    def get_param_names(self):
        names = [r'$\log L_*$', r'$\log \phi_*$']
        if not self.fix_sch_al:
            names.append(r'$\alpha$')
        if not self.fix_comp:
            names += [r'$F_{{\rm 50},%d}$' % (i) for i in range(self.nfields)] + [r'$\alpha_C$']
        return names

    def get_params(self):
        vals = list(self._current_theta()[0])
        self.nfreeparams = len(vals)
        return vals

    def _device_sampler_engine(self, func):
        # user rows equal engine rows only when every parameter of the free-completeness model is sampled
        if not self.fix_comp and not self.fix_sch_al and func == self.lnprob:
            return self._engine('free')
        return None

    def fit_model(self):
        """Run the ensemble sampler on ``lnprob`` (or ``lnprob_fix_comp``) and keep the post-burn-in samples with
        their ln-probabilities in ``self.samples`` (reference lumfuncmcmc.py:479-513)."""
        self._run_sampler(self.lnprob_fix_comp if self.fix_comp else self.lnprob)

    # ------------------------------------------------------------------ 1/V_eff and posterior summaries
    def VeffLF(self):
        """1/V_eff binned luminosity function with bootstrap errors -> ``Lavg``, ``lfbinorig``, ``var``
        (reference lumfuncmcmc.py:515-525)."""
        self.getFlim()
        # the reference evaluates rootsf.ev(self.Flims_arr, self.alpha) for every source (lumfuncmcmc.py:520); Flims_arr is
        # constant within a field and the spline is evaluated point by point, so the K per-field values are the same numbers
        flims = np.asarray(self.Flim, dtype=np.float64)
        self._veff(self.rootsf.ev(flims, np.full(flims.shape, float(self.alpha))))

    def _median_model(self, nsamples, rndsamples):
        Flims, alphas = np.zeros((rndsamples, self.nfields)), np.zeros(rndsamples)
        lstars, lf = np.zeros(rndsamples), []
        for i in np.arange(rndsamples):
            ind = np.random.randint(0, nsamples.shape[0])
            self.set_parameters_from_list(nsamples[ind, :])
            Flims[i], alphas[i], lstars[i] = self.Flim, self.alpha, self.Lstar
            lf.append(TrueLumFunc(self.lum, self.sch_al, self.Lstar, self.phistar))
        self.medianLF = np.median(np.array(lf), axis=0)
        self.Flim, self.alpha = list(np.median(Flims, axis=0)), np.median(alphas)
        return lf, lstars

    def set_median_fit(self, rndsamples=200, lnprobcut=7.5):
        """Median model LF over random posterior draws, median completeness parameters, then ``VeffLF``
        (reference lumfuncmcmc.py:527-567)."""
        nsamples = self._lnprob_selection(lnprobcut, drop_lnprob=False)
        self.log.info("Shape of nsamples (with a lnprobcut applied)")
        self.log.info(nsamples.shape)
        self._median_model(nsamples, rndsamples)
        self.VeffLF()

    def triangle_plot(self, outname, lnprobcut=7.5, imgtype='png'):
        """Corner plot with the LF panel (reference lumfuncmcmc.py:569-651).  The numerical side effects
        (``medianLF``, median completeness, ``VeffLF``) always happen; the figure needs matplotlib + corner."""
        nsamples = self._lnprob_selection(lnprobcut, drop_lnprob=False)
        self.log.info("Shape of nsamples (with a lnprobcut applied)")
        self.log.info(nsamples.shape)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import corner
        except ImportError:
            self.log.info("matplotlib/corner not available: skipping the figure, computing its data products")
            self._median_model(nsamples, 200)
            self.roots_ln = self.rootsf.ev(self.Flim, self.alpha)
            self.VeffLF()
            return
        names = self.get_param_names()
        nd = len(nsamples[0])
        fs = 11 + int(round(0.75 * nd))
        fig = corner.corner(nsamples[:, :-1], labels=names, range=[.95] * len(names), label_kwargs={"fontsize": fs},
                            show_titles=True, title_kwargs={"fontsize": fs - 2}, quantiles=[0.16, 0.5, 0.84], bins=30)
        w = fig.get_figwidth()
        if nd >= 4:
            fig.set_figwidth(w - (nd - 13) * 0.025 * w)
            box = [0.50 - 0.008 * (nd - 4), 0.78 - 0.001 * (nd - 4), 0.48 + 0.008 * (nd - 4), 0.19 + 0.001 * (nd - 4)]
        else:
            box = [0.67, 0.75, 0.32, 0.23]
        ax = fig.add_subplot(3, 1, 1)
        ax.set_position(box)
        ax.set_yscale('log')
        ax.set_xlabel(r"$\log$ L (erg s$^{-1}$)")
        ax.set_ylabel(r"$\phi_{\rm{true}}$ (Mpc$^{-3}$ dex$^{-1}$)")
        ax.minorticks_on()
        order = np.argsort(self.lum)
        lf, lstars = self._median_model(nsamples, 200)
        for model in lf:
            ax.plot(self.lum[order], model[order], color='r', linestyle='solid', alpha=0.1)
        self.roots_ln = self.rootsf.ev(self.Flim, self.alpha)
        self.VeffLF()
        ax.plot(self.lum[order], self.medianLF[order], color='dimgray', linestyle='solid')
        xmin = np.log10(V.get_L_constF(max(self.roots_ln), min(self.z))) if max(self.roots_ln) > 0 else min(self.lum)
        xmax = min(max(self.lum), np.median(lstars) + 1.0)
        ax.set_xlim(left=xmin, right=xmax)
        sel = np.logical_and(self.lum <= xmax, self.lum >= xmin)
        if sel.any():
            ax.set_ylim(bottom=np.min(self.medianLF[sel]), top=np.max(self.medianLF[sel]))
        fig.savefig("%s.%s" % (outname, imgtype), dpi=200)
        plt.close(fig)
