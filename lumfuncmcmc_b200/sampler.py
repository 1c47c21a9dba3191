"""Affine-invariant ensemble sampler (Goodman & Weare stretch move) with the slice of emcee's interface the
reference uses: ``EnsembleSampler(nwalkers, ndim, log_prob_fn)``, ``run_mcmc(pos, nsteps, rstate0=...)``,
``.chain`` (nwalkers, nsteps, ndim), ``.lnprobability`` (nwalkers, nsteps), ``.acceptance_fraction``, ``.acor``
(reference lumfuncmcmc.py:489-513).  emcee itself is not installed in this image and its version is not pinned
by the reference, so this is a from-scratch implementation of the published algorithm, "parity unpinned" against
emcee's random stream: chains agree statistically, not sample by sample.

With ``vectorize=True`` the log-probability function receives a whole half-ensemble (n, ndim) per call -- that
is what hands the GPU engine its batch; the reference's scalar mode (one walker per call) is also supported.
"""
import numpy as np


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, vectorize=False, args=(), kwargs=None,
                 randomize_split=True):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("need an even number of walkers, at least twice the number of dimensions")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn, self.vectorize = log_prob_fn, bool(vectorize)
        self.args, self.kwargs = tuple(args), dict(kwargs or {})
        self.randomize_split = randomize_split
        self._random = np.random.RandomState()
        self.reset()

    def reset(self):
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._lnprob = np.empty((0, self.nwalkers))
        self.naccepted = np.zeros(self.nwalkers)
        self.iterations = 0
        self.ncalls = 0

    # -- evaluation -------------------------------------------------------------------------------
    def _logp(self, pos):
        self.ncalls += 1
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(pos, *self.args, **self.kwargs), dtype=np.float64)
        else:
            lp = np.array([self.log_prob_fn(p, *self.args, **self.kwargs) for p in pos], dtype=np.float64)
        if lp.shape != (len(pos),):
            raise ValueError("log_prob_fn returned shape %s for %d positions" % (lp.shape, len(pos)))
        if np.isnan(lp).any():
            raise ValueError("log_prob_fn returned NaN")
        return lp

    # -- one ensemble update: both halves -----------------------------------------------------------
    def _step(self, pos, lp):
        n, half = self.nwalkers, self.nwalkers // 2
        order = self._random.permutation(n) if self.randomize_split else np.arange(n)
        sets = (order[:half], order[half:])
        for move, other in ((0, 1), (1, 0)):
            s, c = sets[move], sets[other]
            # z ~ g(z) propto 1/sqrt(z) on [1/a, a]
            zz = ((self.a - 1.0) * self._random.rand(len(s)) + 1.0) ** 2 / self.a
            partner = c[self._random.randint(len(c), size=len(s))]
            prop = pos[partner] - (pos[partner] - pos[s]) * zz[:, None]
            new_lp = self._logp(prop)
            with np.errstate(invalid='ignore'):
                lnratio = (self.ndim - 1.0) * np.log(zz) + new_lp - lp[s]
            accept = np.log(self._random.rand(len(s))) < lnratio        # NaN (inf - inf) compares False
            idx = s[accept]
            pos[idx] = prop[accept]
            lp[idx] = new_lp[accept]
            self.naccepted[idx] += 1
        return pos, lp

    def run_mcmc(self, pos0, nsteps, rstate0=None, lnprob0=None, progress=False):
        """Advance the ensemble ``nsteps`` updates from ``pos0``; returns (pos, lnprob, rstate)."""
        if rstate0 is not None:
            self._random.set_state(rstate0)
        pos = np.array(pos0, dtype=np.float64, copy=True)
        if pos.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial positions must have shape (nwalkers, ndim)")
        lp = self._logp(pos) if lnprob0 is None else np.array(lnprob0, dtype=np.float64, copy=True)
        chain = np.empty((nsteps, self.nwalkers, self.ndim))
        lnprob = np.empty((nsteps, self.nwalkers))
        for t in range(nsteps):
            pos, lp = self._step(pos, lp)
            chain[t], lnprob[t] = pos, lp
        self._chain = np.concatenate([self._chain, chain])
        self._lnprob = np.concatenate([self._lnprob, lnprob])
        self.iterations += nsteps
        return pos, lp, self._random.get_state()

    # -- emcee-style views --------------------------------------------------------------------------
    @property
    def chain(self):
        return np.swapaxes(self._chain, 0, 1)

    @property
    def lnprobability(self):
        return self._lnprob.T

    @property
    def flatchain(self):
        return self.chain.reshape(-1, self.ndim)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    def get_autocorr_time(self, c=5.0):
        return integrated_time(self._chain, c=c)

    @property
    def acor(self):
        return self.get_autocorr_time()


def _autocorr_1d(x):
    n = 1 << int(np.ceil(np.log2(max(len(x), 2))))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0] if acf[0] > 0 else np.ones_like(acf)


def integrated_time(chain, c=5.0):
    """Integrated autocorrelation time per dimension with Sokal's automatic window (M >= c tau); ``chain`` has
    shape (nsteps, nwalkers, ndim) and the autocorrelation function is averaged over walkers.  Never raises on
    short chains: it returns the (then unreliable) estimate, which the caller clips (lumfuncmcmc.py:499-501)."""
    chain = np.asarray(chain, dtype=np.float64)
    nsteps, nwalkers, ndim = chain.shape
    tau = np.empty(ndim)
    for d in range(ndim):
        acf = np.zeros(nsteps)
        for w in range(nwalkers):
            acf += _autocorr_1d(chain[:, w, d])
        acf /= nwalkers
        taus = 2.0 * np.cumsum(acf) - 1.0
        window = np.arange(nsteps) >= c * taus
        m = np.argmax(window) if window.any() else nsteps - 1
        tau[d] = taus[m]
    return tau
