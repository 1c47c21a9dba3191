"""Python face of the CUDA likelihood engine: NumPy in, NumPy out (or torch device tensors, zero-copy).

``LikelihoodEngine(inputs, kind)`` uploads the arrays the reference's ``lnlike`` reads from ``self``
(reference lumfuncmcmc.py:360-393, lumfuncmcmc_z.py:364-376; produced once by the set-up chain
lumfuncmcmc.py:180-235) and then evaluates ``lnprob`` for a whole walker ensemble per call.

``inputs`` keys (float64 arrays unless noted) -- the same dict the oracle consumes:
  lum, z, zint, DLarr, field_ind (int64, K+1), Omega_0 (K), Flim (K), alpha, fcmin, logL (K,S,S), zarr (S),
  DL_zarr (S), volume_part (S), [Om_arr (N), integ_part (K,S,S)], the prior boxes ``*_lims``, ``sch_al``,
  ``fix_sch_al`` and (kind 'z') the pivots z1, z2, z3.
"""
import ctypes as C

import numpy as np

from . import _lib

KINDS = {'free': _lib.LF_MODEL_FREE, 'fixed': _lib.LF_MODEL_FIXED, 'z': _lib.LF_MODEL_Z}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def source_flux(inp):
    """Per-source flux as the reference's ``Omega`` derives it from ``lum`` and the *interpolated* D_L
    (reference lumfuncmcmc.py:69-70 with dLzfunc = interp1d(zint, DLarr), :196) -- walker-independent, so it is
    evaluated once here instead of once per likelihood call."""
    z = np.asarray(inp['z'], dtype=np.float64)
    zint, DLarr = np.asarray(inp['zint']), np.asarray(inp['DLarr'])
    if z.size and (z.min() < zint[0] or z.max() > zint[-1]):
        raise ValueError("source redshift outside the D_L interpolation table")
    # the linear interpolant the reference builds (lumfuncmcmc.py:196: interp1d -> numpy.interp), evaluated once; for a
    # large catalogue on the GPU, bit-identical (lf_interp_linear)
    from .setup_gpu import LinearTable
    DL = LinearTable(zint, DLarr)(z)
    L = 10 ** np.asarray(inp['lum'], dtype=np.float64)
    return L / (4.0 * np.pi * (3.086e24 * DL) ** 2)


class _VeffOps:
    """1/V_eff weights, binned LF and bootstrap replicates on a context (``self._ctx`` / ``self.lib``)."""

    _phi_gen = 0             # bumped whenever the weights resident on the device are replaced
    _sample_owner = None     # whoever uploaded the resident sample (lfbase.LFBase._veff)

    def veff_bin(self, flux, lum, field_ind, flim, alpha, fcmin, sum_omega, vol_int, edges, vol_per_source=None,
                 valid=None, want_phi=True):
        """1/V_eff weights and the binned LF of the original sample (reference lumfuncmcmc.py:515-525,
        VmaxLumFunc.py:336-350).  Returns (phi or None, counts int64[nbins], sumphi float64[nbins])."""
        flux, lum, edges, flim = _f64(flux), _f64(lum), _f64(edges), _f64(flim)
        fi = np.ascontiguousarray(field_ind, dtype=np.int64)
        n, nb = flux.shape[0], edges.shape[0] - 1
        vps = None if vol_per_source is None else _f64(vol_per_source)
        val = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        phi = np.empty(n, dtype=np.float64) if want_phi else None
        self._veff_nbins = nb
        self._phi_gen += 1
        self._sample_owner = None
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        _lib.check(self.lib.lf_veff_bin(self._ctx, n, _ptr(flux), _ptr(lum), _ptr(fi), len(flim), _ptr(flim),
                                        float(alpha), float(fcmin) if fcmin else 0.0, float(sum_omega), float(vol_int),
                                        _ptr(vps), _ptr(val), _ptr(edges), nb, _ptr(phi), _ptr(counts), _ptr(sums)),
                   self.lib)
        return phi, counts, sums

    # ---- sample resident across calls (VeffLF runs once per posterior summary on the same catalogue) ----
    def veff_set_sample(self, flux, lum, field_ind):
        """Upload flux (cgs), log-luminosity and the field offsets once (``lf_veff_set_sample``)."""
        flux, lum = _f64(flux), _f64(lum)
        fi = np.ascontiguousarray(field_ind, dtype=np.int64)
        _lib.check(self.lib.lf_veff_set_sample(self._ctx, flux.shape[0], _ptr(flux), _ptr(lum), _ptr(fi), len(fi) - 1), self.lib)
        self._veff_n = int(flux.shape[0])

    def veff_bin_resident(self, flim, alpha, fcmin, sum_omega, vol_int, edges, device_volumes=False, want_phi=False):
        """Weights + binned LF of the resident sample; the weights stay on the device (:meth:`veff_phi` fetches them).
        Returns (phi or None, counts, sumphi)."""
        flim, edges = _f64(flim), _f64(edges)
        nb = edges.shape[0] - 1
        phi = np.empty(self._veff_n, dtype=np.float64) if want_phi else None
        self._veff_nbins = nb
        self._phi_gen += 1
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        _lib.check(self.lib.lf_veff_bin_resident(self._ctx, _ptr(flim), float(alpha), float(fcmin) if fcmin else 0.0,
                                                 float(sum_omega), float(vol_int), int(bool(device_volumes)), _ptr(edges), nb,
                                                 _ptr(phi), _ptr(counts), _ptr(sums)), self.lib)
        return phi, counts, sums

    def veff_phi(self):
        """The resident per-source weights as a host array (``lf_veff_get_phi``)."""
        phi = np.empty(self._veff_n, dtype=np.float64)
        _lib.check(self.lib.lf_veff_get_phi(self._ctx, _ptr(phi)), self.lib)
        return phi

    def veff_set_volume_table(self, cosmo, zk, dVk):
        """Cosmology + knots of the dV/dz interpolant for :meth:`veff_volumes` (``lf_veff_set_volume_table``)."""
        from .setup_gpu import cosmology_struct
        zk, dVk = _f64(zk), _f64(dVk)
        c, cum = cosmology_struct(cosmo, float(zk[-1]))
        _lib.check(self.lib.lf_veff_set_volume_table(self._ctx, C.byref(c), _ptr(cum), cum.shape[0], zk.shape[0], _ptr(zk),
                                                     _ptr(dVk)), self.lib)

    def veff_volumes(self, zmin, zmax, DL_zmin, DL_zmax, fmin, want=False):
        """Per-source upper redshift limit, volume and validity for min_comp_frac > 0 on the resident sample; the results
        stay on the device for ``veff_bin_resident(device_volumes=True)``.  ``want=True`` also returns (zmax_i, vol_i, valid_i)."""
        fmin = _f64(fmin)
        n = self._veff_n
        zm = np.empty(n) if want else None
        vol = np.empty(n) if want else None
        val = np.empty(n, dtype=np.uint8) if want else None
        _lib.check(self.lib.lf_veff_volumes(self._ctx, float(zmin), float(zmax), float(DL_zmin), float(DL_zmax), _ptr(fmin),
                                            _ptr(zm), _ptr(vol), _ptr(val)), self.lib)
        return (zm, vol, val) if want else None

    def boot_bin(self, mult):
        """One bootstrap replicate on the sample left resident by :meth:`veff_bin`."""
        mult = np.ascontiguousarray(mult, dtype=np.int32)
        nb = self._nbins_resident()
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        _lib.check(self.lib.lf_boot_bin(self._ctx, _ptr(mult), _ptr(counts), _ptr(sums)), self.lib)
        return counts, sums

    def boot_mt_set_state(self, state):
        """Hand NumPy's legacy generator state (``np.random.get_state()``) to the device MT19937 stream."""
        if state[0] != 'MT19937':
            raise ValueError("the device stream reproduces NumPy's legacy MT19937 generator only")
        key = np.ascontiguousarray(state[1], dtype=np.uint32)
        _lib.check(self.lib.lf_boot_mt_set_state(self._ctx, _ptr(key), int(state[2])), self.lib)
        self._mt_tail = tuple(state[3:])

    def boot_mt_get_state(self):
        """The generator state after the replicates drawn on the device, in ``np.random.set_state`` form."""
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int32()
        _lib.check(self.lib.lf_boot_mt_get_state(self._ctx, _ptr(key), C.byref(pos)), self.lib)
        return ('MT19937', key, int(pos.value)) + self._mt_tail

    def boot_bin_mt(self):
        """One bootstrap replicate whose indices are NumPy's ``np.random.randint(n, size=n)`` stream, drawn on the device."""
        nb = self._nbins_resident()
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        _lib.check(self.lib.lf_boot_bin_mt(self._ctx, _ptr(counts), _ptr(sums)), self.lib)
        return counts, sums

    def boot_bin_device(self, seed, replicate):
        """One bootstrap replicate resampled ON the device (Philox stream keyed by ``seed``; counter = draw, replicate)."""
        nb = self._nbins_resident()
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        _lib.check(self.lib.lf_boot_bin_device(self._ctx, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(replicate),
                                               _ptr(counts), _ptr(sums)), self.lib)
        return counts, sums

    def _nbins_resident(self):
        return getattr(self, '_veff_nbins', 0)

    def last_kernel_ms(self):
        ms = C.c_double()
        _lib.check(self.lib.lf_last_kernel_ms(self._ctx, C.byref(ms)), self.lib)
        return ms.value

    def bin_weights(self, lum, phi, edges):
        """Counts and sum of caller-provided weights per half-open bin (reference VmaxLumFunc.py:345-350).
        ``lum is None and phi is None``: bin the sample and weights already resident on the device."""
        edges = _f64(edges)
        nb = edges.shape[0] - 1
        self._veff_nbins = nb
        counts, sums = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.float64)
        if lum is None and phi is None:
            _lib.check(self.lib.lf_bin_weights(self._ctx, 0, None, None, _ptr(edges), nb, _ptr(counts), _ptr(sums)), self.lib)
            return counts, sums
        lum, phi = _f64(lum), _f64(phi)
        self._phi_gen += 1
        self._sample_owner = None
        _lib.check(self.lib.lf_bin_weights(self._ctx, lum.shape[0], _ptr(lum), _ptr(phi), _ptr(edges), nb,
                                           _ptr(counts), _ptr(sums)), self.lib)
        return counts, sums

    def close(self):
        if getattr(self, '_ctx', None) is not None and self._ctx.value:
            self.lib.lf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VeffEngine(_VeffOps):
    """A bare context (no likelihood inputs) for the 1/V_eff kernels."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        cfg = _lib.LFConfig()
        cfg.model, cfg.precision, cfg.device, cfg.nfields, cfg.size_ln = _lib.LF_MODEL_FREE, _lib.LF_PREC_F64, int(device), 1, 2
        self._ctx = C.c_void_p()
        _lib.check(self.lib.lf_create(C.byref(self._ctx), C.byref(cfg)), self.lib)
        self.device = int(device)


class LikelihoodEngine(_VeffOps):
    """One engine context on one GPU holding one shard of sources."""

    def __init__(self, inp, kind, device=0, force_literal=False, quadrature_share=(0, 1), precision='f64', compress=False):
        if precision not in ('f64', 'f32'):
            raise ValueError("precision must be 'f64' or 'f32'")
        self.precision = precision
        if kind not in KINDS:
            raise ValueError("kind must be one of %s" % sorted(KINDS))
        self.kind = kind
        self.lib = _lib.load()
        K = len(inp['Flim'])
        if K > _lib.LF_MAX_FIELDS:
            raise ValueError("at most %d fields" % _lib.LF_MAX_FIELDS)
        S = int(np.asarray(inp['zarr']).shape[0])
        fix_sch_al = bool(inp.get('fix_sch_al', False))
        cfg = _lib.LFConfig()
        cfg.model, cfg.device = KINDS[kind], int(device)
        cfg.precision = _lib.LF_PREC_F32 if precision == 'f32' else _lib.LF_PREC_F64
        cfg.nfields, cfg.size_ln, cfg.fix_sch_al = K, S, int(fix_sch_al)
        cfg.force_literal = int(bool(force_literal))
        cfg.fcmin = float(inp['fcmin']) if inp['fcmin'] else 0.0
        cfg.sch_al = float(inp['sch_al'])

        def box(name, default):
            lims = inp.get(name, default)
            return (C.c_double * 2)(float(lims[0]), float(lims[1]))

        cfg.Lstar_lims, cfg.phistar_lims = box('Lstar_lims', (40.0, 45.0)), box('phistar_lims', (-8.0, 5.0))
        cfg.sch_al_lims = box('sch_al_lims', (-3.0, 1.0))
        cfg.Flim_lims, cfg.alpha_lims = box('Flim_lims', (1.0, 6.0)), box('alpha_lims', (1.0, 7.0))
        cfg.z_pivots = (C.c_double * 3)(float(inp.get('z1', 1.20)), float(inp.get('z2', 1.53)), float(inp.get('z3', 1.86)))
        # the reference's single-z lnprior range-checks the parameters that are NOT sampled as well
        # (lumfuncmcmc.py:347-354); they are constants here, so the check is done once
        ok = True
        if kind != 'z':
            if fix_sch_al:
                ok &= cfg.sch_al_lims[0] <= cfg.sch_al <= cfg.sch_al_lims[1]
            if kind == 'fixed':
                ok &= all(cfg.Flim_lims[0] <= f <= cfg.Flim_lims[1] for f in np.asarray(inp['Flim'], dtype=np.float64))
                ok &= cfg.alpha_lims[0] <= float(inp['alpha']) <= cfg.alpha_lims[1]
        cfg.fixed_prior_ok = int(bool(ok))
        self._ctx = C.c_void_p()
        _lib.check(self.lib.lf_create(C.byref(self._ctx), C.byref(cfg)), self.lib)
        self.ndim = self.lib.lf_ndim(self._ctx)
        self.device = int(device)
        self.nfields, self.size_ln = K, S

        # ---- sources -------------------------------------------------------------------------
        lum = _f64(inp['lum'])
        fi = np.ascontiguousarray(inp['field_ind'], dtype=np.int64)
        self.nsources = int(lum.shape[0])
        flux = zz = om = om0i = None
        if kind == 'free':
            flux = _f64(inp['flux_src']) if 'flux_src' in inp else _f64(source_flux(inp))
            # dtype=int copy of the areas: truncation toward zero (lumfuncmcmc.py:285)
            om0i = np.ascontiguousarray(np.asarray(inp['Omega_0'], dtype=np.float64).astype(np.int64))
        else:
            om = _f64(inp['Om_arr'])
            if kind == 'z':
                zz = _f64(inp['z'])
        _lib.check(self.lib.lf_set_sources(self._ctx, self.nsources, _ptr(lum), _ptr(flux), _ptr(zz), _ptr(om),
                                           _ptr(fi), _ptr(om0i)), self.lib)
        self._flux_host = flux                      # kept for compress_catalogue (free model)
        # ---- quadrature grid -----------------------------------------------------------------
        logL = _f64(inp['logL'])
        if logL.shape != (K, S, S):
            raise ValueError("logL must have shape (K, S, S)")
        zarr = _f64(inp['zarr'])
        if kind == 'free':
            dlz, vol, om0 = _f64(inp['DL_zarr']), _f64(inp['volume_part']), _f64(inp['Omega_0'])
            _lib.check(self.lib.lf_set_grid(self._ctx, _ptr(logL), _ptr(zarr), _ptr(dlz), _ptr(vol), None, _ptr(om0)),
                       self.lib)
        else:
            ip = _f64(inp['integ_part'])
            if ip.shape != (K, S, S):
                raise ValueError("integ_part must have shape (K, S, S)")
            _lib.check(self.lib.lf_set_grid(self._ctx, _ptr(logL), _ptr(zarr), None, None, _ptr(ip), None), self.lib)
        if tuple(quadrature_share) != (0, 1):
            self.set_quadrature_share(*quadrature_share)
        self.npseudo = 0
        self._z_host, self._lum_host, self._fi_host = zz, lum, fi
        self._z_pivots = tuple(cfg.z_pivots)
        self._Lstar_lims = tuple(cfg.Lstar_lims)
        if compress:
            if kind == 'fixed' or precision != 'f64':
                raise ValueError("compress=True applies to the free-completeness and z-evolving models in FP64")
            opts = compress if isinstance(compress, dict) else {}
            if kind == 'free':
                self.compress_catalogue(flux, fi, alpha_max=float(cfg.alpha_lims[1]), **opts)
            else:
                self.compress_catalogue_z(**opts)

    def compress_catalogue(self, flux, field_ind, alpha_max, nodes=12, bin_dex=None):
        """Switch the fast kernels' source sum to the weighted pseudo-source form (see :mod:`.compress`)."""
        from .compress import compress_sources
        xi, w, cfi = compress_sources(np.log10(flux), field_ind, alpha_max, nodes=nodes, bin_dex=bin_dex)
        xi, w = _f64(xi), _f64(w)
        _lib.check(self.lib.lf_set_compressed_sources(self._ctx, xi.shape[0], _ptr(xi), _ptr(w), _ptr(cfi), float(alpha_max)),
                   self.lib)
        self.npseudo = int(xi.shape[0])
        return self.npseudo

    def compress_catalogue_z(self, slope_max=None, nodes=12, bin_z=None):
        """z-evolving model: weighted pseudo-sources in redshift (see :func:`.compress.compress_sources_z`).  Walkers
        whose |dL*/dz| exceeds ``slope_max`` over the catalogue's redshift range take the literal kernels."""
        from .compress import compress_sources_z
        if slope_max is None:                       # four times the steepest secant the prior box allows between pivots
            gap = min(self._z_pivots[1] - self._z_pivots[0], self._z_pivots[2] - self._z_pivots[1])
            slope_max = 4.0 * (self._Lstar_lims[1] - self._Lstar_lims[0]) / gap
        xi, v, cfi = compress_sources_z(self._z_host, self._lum_host, self._fi_host, slope_max, nodes=nodes, bin_z=bin_z)
        xi, v = _f64(xi), _f64(v)
        _lib.check(self.lib.lf_set_compressed_sources(self._ctx, xi.shape[0], _ptr(xi), _ptr(v), _ptr(cfi), float(slope_max)),
                   self.lib)
        self.npseudo = int(xi.shape[0])
        return self.npseudo

    def uncompress_catalogue(self):
        _lib.check(self.lib.lf_set_compressed_sources(self._ctx, 0, None, None, None, 1.0), self.lib)
        self.npseudo = 0

    # ------------------------------------------------------------------------------------------
    def set_quadrature_share(self, share, nshare):
        _lib.check(self.lib.lf_set_quadrature_share(self._ctx, int(share), int(nshare)), self.lib)

    def lnlike(self, thetas):
        """Like :meth:`lnprob` but without the prior gate (the reference's ``lnlike`` never consults the prior)."""
        _lib.check(self.lib.lf_set_prior_gate(self._ctx, 0), self.lib)
        try:
            return self.lnprob(thetas)
        finally:
            _lib.check(self.lib.lf_set_prior_gate(self._ctx, 1), self.lib)

    def lnprob(self, thetas):
        """thetas: (ndim,) or (W, ndim) host array -> float or (W,) array.  H2D, kernels, D2H, sync."""
        th = np.asarray(thetas, dtype=np.float64)
        scalar = th.ndim == 1
        th = np.ascontiguousarray(np.atleast_2d(th))
        if th.shape[1] != self.ndim:
            raise ValueError("theta has %d parameters, this model takes %d" % (th.shape[1], self.ndim))
        out = np.empty(th.shape[0], dtype=np.float64)
        _lib.check(self.lib.lf_lnprob_batch(self._ctx, _ptr(th), th.shape[0], _ptr(out)), self.lib)
        return float(out[0]) if scalar else out

    def lnprob_device(self, d_thetas, d_out=None, stream=None):
        """torch CUDA float64 tensors, asynchronous on ``stream`` (default: torch's current stream)."""
        import torch
        if d_thetas.dtype != torch.float64 or not d_thetas.is_cuda or not d_thetas.is_contiguous():
            raise ValueError("d_thetas must be a contiguous CUDA float64 tensor")
        W = d_thetas.shape[0]
        if d_thetas.dim() != 2 or d_thetas.shape[1] != self.ndim:
            raise ValueError("d_thetas must have shape (W, %d)" % self.ndim)
        if d_out is None:
            d_out = torch.empty(W, dtype=torch.float64, device=d_thetas.device)
        st = torch.cuda.current_stream(d_thetas.device) if stream is None else stream
        _lib.check(self.lib.lf_lnprob_batch_device(self._ctx, C.c_void_p(d_thetas.data_ptr()), W,
                                                   C.c_void_p(d_out.data_ptr()), C.c_void_p(st.cuda_stream)), self.lib)
        return d_out

    def sampler_run(self, pos0, nsteps, seed, a=2.0, step0=0, store_chain=True):
        """Device-resident stretch-move run (see include/lf_engine.h, lf_sampler_run).  Returns a dict with
        chain (nsteps, W, ndim), lnprob (nsteps, W), naccepted (W,), pos (W, ndim), lp (W,), device_ms."""
        pos0 = np.ascontiguousarray(pos0, dtype=np.float64)
        W, nd = pos0.shape
        if nd != self.ndim:
            raise ValueError("positions have %d parameters, this model takes %d" % (nd, self.ndim))
        nsteps = int(nsteps)
        chain = np.empty((nsteps, W, nd), dtype=np.float64) if store_chain else None
        lnp = np.empty((nsteps, W), dtype=np.float64) if store_chain else None
        nacc = np.zeros(W, dtype=np.int64)
        pos, lp = np.empty((W, nd), dtype=np.float64), np.empty(W, dtype=np.float64)
        _lib.check(self.lib.lf_sampler_run(self._ctx, _ptr(pos0), W, nsteps, C.c_uint64(int(seed) & (2 ** 64 - 1)), float(a),
                                           int(step0), _ptr(chain), _ptr(lnp), _ptr(nacc), _ptr(pos), _ptr(lp)), self.lib)
        ms = C.c_double()
        _lib.check(self.lib.lf_sampler_last_ms(self._ctx, C.byref(ms)), self.lib)
        return dict(chain=chain, lnprob=lnp, naccepted=nacc, pos=pos, lp=lp, device_ms=ms.value)

    def set_walker_sharding(self, enabled=True):
        """Several ranks, every rank holding all sources: ``sampler_run`` evaluates this rank's slice of each half-ensemble
        and the peer-memory exchange gathers the slices (``lf_set_walker_sharding``)."""
        _lib.check(self.lib.lf_set_walker_sharding(self._ctx, int(bool(enabled))), self.lib)

    # ---- peer-memory exchange (multi-GPU without NCCL on the data path) ----
    def peer_buffer_create(self, rank, world, wcap):
        """Allocate this rank's receive buffer; returns its 64-byte CUDA IPC handle (bytes)."""
        h = (C.c_ubyte * 64)()
        _lib.check(self.lib.lf_peer_buffer_create(self._ctx, int(rank), int(world), int(wcap), h), self.lib)
        return bytes(h)

    def peer_buffer_connect(self, handles):
        """``handles``: list of the world ranks' IPC handles, in rank order."""
        blob = b''.join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.lf_peer_buffer_connect(self._ctx, buf), self.lib)

    def allreduce_device(self, d_vec, stream=None):
        """In-place SUM over ranks of a CUDA float64 vector, one kernel over peer memory (asynchronous)."""
        import torch
        st = torch.cuda.current_stream(d_vec.device) if stream is None else stream
        _lib.check(self.lib.lf_allreduce_device(self._ctx, C.c_void_p(d_vec.data_ptr()), d_vec.shape[0],
                                                C.c_void_p(st.cuda_stream)), self.lib)
        return d_vec

    def peer_timed_out(self):
        """True once a wait for another rank has expired (sticky until :meth:`peer_reset`; results since then are NaN)."""
        v = C.c_int32()
        _lib.check(self.lib.lf_peer_status(self._ctx, C.byref(v)), self.lib)
        return bool(v.value)

    def peer_reset(self):
        _lib.check(self.lib.lf_peer_reset(self._ctx), self.lib)

    def peer_set_timeout(self, seconds):
        _lib.check(self.lib.lf_peer_set_timeout(self._ctx, float(seconds)), self.lib)

    def last_call_info(self):
        counts = (C.c_int64 * 3)()
        launches = C.c_int64()
        _lib.check(self.lib.lf_last_call_info(self._ctx, counts, C.byref(launches)), self.lib)
        return dict(rejected=counts[0], fast=counts[1], literal=counts[2], launches=launches.value)

    def mufu_peak(self, iters=20000):
        """Measured MUFU ex2.approx.f32 rate of this GPU (thread-instructions per second) and the run time in ms."""
        rate, ms = C.c_double(), C.c_double()
        _lib.check(self.lib.lf_mufu_peak(self._ctx, int(iters), C.byref(rate), C.byref(ms)), self.lib)
        return rate.value, ms.value

    def fp64_peak(self, iters=20000):
        """Measured register-only DFMA rate of this GPU (thread-instructions per second) and the run time in ms."""
        rate, ms = C.c_double(), C.c_double()
        _lib.check(self.lib.lf_fp64_peak(self._ctx, int(iters), C.byref(rate), C.byref(ms)), self.lib)
        return rate.value, ms.value
