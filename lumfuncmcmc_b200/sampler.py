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


class DeviceEnsembleSampler(EnsembleSampler):
    """Same interface, but the whole run lives on the GPU (``LikelihoodEngine.sampler_run``): proposals from a
    Philox4x32-10 counter stream, log-posterior from the engine, accept/reject and the chain all on the device, one
    CUDA graph per ensemble update.  The split of the ensemble is the fixed first-half / second-half split of
    emcee 2.x.  ``rstate0`` only seeds the counter stream (its first state word); chains are reproducible for a given
    seed and equal, sample for sample, to :func:`philox_stretch_reference` driven by the same log-posterior."""

    def __init__(self, nwalkers, ndim, engine, a=2.0, seed=None):
        super().__init__(nwalkers, ndim, engine.lnprob, a=a, vectorize=True, randomize_split=False)
        self.engine = engine
        self.seed = seed
        self.device_ms = 0.0

    def run_mcmc(self, pos0, nsteps, rstate0=None, lnprob0=None, progress=False, seed=None):
        if seed is None:
            seed = self.seed
        if seed is None:
            if rstate0 is not None:
                seed = int(np.asarray(rstate0[1], dtype=np.uint64)[0]) | (int(np.asarray(rstate0[1], dtype=np.uint64)[1]) << 32)
            else:
                seed = int(self._random.randint(0, 2 ** 31 - 1))
        self.seed = seed
        pos0 = np.array(pos0, dtype=np.float64, copy=True)
        if pos0.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial positions must have shape (nwalkers, ndim)")
        out = self.engine.sampler_run(pos0, nsteps, seed, a=self.a, step0=self.iterations)
        self._chain = np.concatenate([self._chain, out['chain']])
        self._lnprob = np.concatenate([self._lnprob, out['lnprob']])
        self.naccepted += out['naccepted']
        self.iterations += int(nsteps)
        self.ncalls += 2 * int(nsteps) + 2
        self.device_ms += out['device_ms']
        return out['pos'], out['lp'], rstate0


# ---- host replay of the device sampler's random stream (test infrastructure and documentation of the algorithm) ----
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 NumPy arrays (Salmon et al. 2011); returns four uint32 arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox_stretch_reference(log_prob_fn, pos0, nsteps, seed, a=2.0, step0=0):
    """The device sampler's algorithm, operation for operation, on the host: returns (chain, lnprob, naccepted)."""
    pos = np.array(pos0, dtype=np.float64, copy=True)
    W, ndim = pos.shape
    half = W // 2
    lp = np.asarray(log_prob_fn(pos), dtype=np.float64).copy()
    chain, lnp, nacc = np.empty((nsteps, W, ndim)), np.empty((nsteps, W)), np.zeros(W, dtype=np.int64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for t in range(nsteps):
        step = step0 + t
        for h in (0, 1):
            me = h * half + np.arange(half)
            r = philox4x32_10(me, np.full(half, step & 0xFFFFFFFF), np.full(half, step >> 32), np.full(half, h), k0, k1)
            u1 = (r[0].astype(np.float64) + 0.5) * 2.3283064365386963e-10
            u3 = (r[2].astype(np.float64) + 0.5) * 2.3283064365386963e-10
            tt = (a - 1.0) * u1 + 1.0
            z = (tt * tt) / a
            partner = (1 - h) * half + ((r[1].astype(np.uint64) * np.uint64(half)) >> np.uint64(32)).astype(np.int64)
            prop = pos[partner] - (pos[partner] - pos[me]) * z[:, None]
            new_lp = np.asarray(log_prob_fn(prop), dtype=np.float64)
            with np.errstate(invalid='ignore'):
                acc = np.log(u3) < (ndim - 1.0) * np.log(z) + new_lp - lp[me]
            pos[me[acc]] = prop[acc]
            lp[me[acc]] = new_lp[acc]
            nacc[me[acc]] += 1
        chain[t], lnp[t] = pos, lp
    return chain, lnp, nacc


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
