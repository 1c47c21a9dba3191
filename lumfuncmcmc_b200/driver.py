"""Shared body of the two command-line drivers (``run_lumfuncmcmc.py``, ``run_lumfuncmcmc_z.py``).

Same flags, same fallback to ``configLF`` for anything left unset, same catalogue format and the same output files
as the reference's drivers (reference run_lumfuncmcmc.py:39-330, run_lumfuncmcmc_z.py:38-303); astropy's ``Table``
is replaced by ``lumfuncmcmc_b200.tableio.Table``.  One extra flag: ``--device``.
"""
import argparse as ap
import logging
import os
import os.path as op
import sys

import numpy as np
from scipy.optimize import fsolve

from . import VmaxLumFunc as V
from . import configLF
from .tableio import Table

CONFIG_FALLBACK = ['nwalkers', 'nsteps', 'nbins', 'nboot', 'Flim', 'alpha', 'line_name', 'line_plot_name', 'Omega_0',
                   'sch_al', 'sch_al_lims', 'Lstar', 'Lstar_lims', 'phistar', 'phistar_lims', 'Lc', 'Lh',
                   'min_comp_frac', 'param_percentiles', 'output_dict', 'fcmin']


def setup_logging(name):
    log = logging.getLogger(name)
    if not len(log.handlers):
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('[%(levelname)s - %(asctime)s] %(message)s'))
        handler.setLevel(logging.INFO)
        log.setLevel(logging.DEBUG)
        log.addHandler(handler)
    return log


def parse_args(argv=None, evolving=False):
    """Command line first, ``configLF`` for everything that is None or 0 (reference run_lumfuncmcmc.py:39-134)."""
    p = ap.ArgumentParser(description="LumFuncMCMCz" if evolving else "LumFuncMCMC", formatter_class=ap.RawTextHelpFormatter)
    p.add_argument("-f", "--filename", type=str, default=None, help="File to be read for galaxy data")
    p.add_argument("-o", "--output_filename", type=str, default='test.dat', help="Output filename for given run")
    p.add_argument("-nw", "--nwalkers", type=int, default=None, help="Number of walkers")
    p.add_argument("-ns", "--nsteps", type=int, default=None, help="Number of steps")
    p.add_argument("-nbins", "--nbins", type=int, default=None, help="Number of bins of the V_eff luminosity function")
    p.add_argument("-nboot", "--nboot", type=int, default=None, help="Number of bootstrap iterations for V_eff method")
    p.add_argument("-mcf", "--min_comp_frac", type=float, default=None, help="Minimum completeness fraction considered")
    p.add_argument("-al", "--alpha", type=float, default=None, help="Completeness slope")
    p.add_argument('-fl', '--Flim', type=float, nargs='*', default=None, help="Flim for each of the fields")
    p.add_argument("-sa", "--sch_al", type=float, default=None, help="Schechter alpha")
    p.add_argument("-fsa", "--fix_sch_al", action='count', default=0, help="Fix Schechter alpha")
    if not evolving:
        p.add_argument("-fc", "--fix_comp", action='count', default=0, help="Fix completeness")
        p.add_argument("-sr", "--same_rand", action='count', default=0, help="Same random starting point")
    p.add_argument("-ln", "--line_name", type=str, default=None, help="Name of line or band for LF measurement")
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    args = p.parse_args(args=argv)
    args.log = setup_logging('lumfuncmcmc_z' if evolving else 'lumfuncmcmc')
    fallback = CONFIG_FALLBACK + ([] if evolving else ['Flim_lims', 'alpha_lims'])
    for name in fallback:
        if getattr(args, name, None) in [None, 0]:
            setattr(args, name, getattr(configLF, name))
    if args.line_name == 'OIII':
        args.line_plot_name = r'[OIII] $\lambda 5007$'
    if args.line_name == 'Ha':
        args.line_plot_name = r'${\rm{H\alpha}}$'
    if evolving:
        # pivot redshifts by catalogue / line (reference run_lumfuncmcmc_z.py:123-128)
        args.z1, args.z2, args.z3 = (1.20, 1.76, 2.32) if args.filename == 'OIIIGautamWillNoAGN.dat' else (1.20, 1.53, 1.86)
        if args.line_name == 'Ha':
            args.z1, args.z2, args.z3 = 1.18, 1.36, 1.54
    return args


def read_input_file(args):
    """Catalogue -> per-field lists.  Columns: ``Field z ID <line>_flux <line>_flux_e`` (fluxes in 1e-17 cgs) or
    ``<line>_lum [<line>_lum_e]``; sources below the per-field minimum flux are dropped and the survivors are grouped
    by field with cumulative offsets ``field_ind`` (reference run_lumfuncmcmc.py:136-228)."""
    dat = Table.read(args.filename, format='ascii')
    fields, zfull = dat['Field'], dat['z']
    field_names = np.unique(fields)
    if abs(args.min_comp_frac - 0.0) < 1.0e-6:
        roots = np.zeros(len(field_names))
    else:
        roots = np.array([fsolve(lambda x: V.fleming(x, args.Flim[i], args.alpha, args.fcmin) - args.min_comp_frac,
                                 [args.Flim[i]])[0] for i in range(len(field_names))])
    fcol, lcol = '%s_flux' % args.line_name, '%s_lum' % args.line_name
    flux = flux_e = lum = lum_e = None
    field_ind, z = [0], []
    if fcol in dat.columns:
        fl, fe = dat[fcol], dat[fcol + '_e']
        flux, flux_e = [], []
        for i, f in enumerate(field_names):
            keep = np.logical_and(fields == f, fl > roots[i])
            flux.append(fl[keep])
            flux_e.append(fe[keep])
            z.append(zfull[keep])
            field_ind.append(field_ind[-1] + int(keep.sum()))
    elif lcol in dat.columns:
        ll = dat[lcol]
        le = dat[lcol + '_e'] if (lcol + '_e') in dat.columns else None
        lum, lum_e = [], ([] if le is not None else None)
        for f in field_names:
            keep = np.logical_and(fields == f, ll > 0)
            lum.append(ll[keep])
            if le is not None:
                lum_e.append(le[keep])
            z.append(zfull[keep])
            field_ind.append(field_ind[-1] + int(keep.sum()))
    else:
        raise KeyError("catalogue needs a %s or %s column" % (fcol, lcol))
    return z, flux, flux_e, lum, lum_e, field_names, np.array(field_ind)


def _tag(args):
    return '%s_nb%d_nw%d_ns%d_mcf%d' % (args.output_filename.split('.')[0], args.nbins, args.nwalkers, args.nsteps,
                                        int(100 * args.min_comp_frac))


def run(argv, evolving, script_name):
    outdir = 'LFMCMCzOut' if evolving else 'LFMCMCOut'
    os.makedirs(outdir, exist_ok=True)
    if argv is None:
        argv = [a for a in sys.argv[1:]] if op.basename(sys.argv[0]) == script_name else \
            [a for a in sys.argv if a != script_name]
    args = parse_args(argv, evolving)
    z, flux, flux_e, lum, lum_e, field_names, field_ind = read_input_file(args)
    print("Read Input File")
    common = dict(flux=flux, flux_e=flux_e, lum=lum, lum_e=lum_e, Flim=args.Flim, alpha=args.alpha,
                  line_name=args.line_name, line_plot_name=args.line_plot_name, Omega_0=args.Omega_0, nbins=args.nbins,
                  nboot=args.nboot, sch_al=args.sch_al, sch_al_lims=args.sch_al_lims, Lstar=args.Lstar,
                  Lstar_lims=args.Lstar_lims, phistar=args.phistar, phistar_lims=args.phistar_lims, Lc=args.Lc,
                  Lh=args.Lh, nwalkers=args.nwalkers, nsteps=args.nsteps, min_comp_frac=args.min_comp_frac,
                  field_names=field_names, field_ind=field_ind, fix_sch_al=args.fix_sch_al, device=args.device)
    if evolving:
        from .lumfuncmcmc_z import LumFuncMCMCz
        LFmod = LumFuncMCMCz(z, z1=args.z1, z2=args.z2, z3=args.z3, **common)
        print("Initialized LumFuncMCMCz class")
    else:
        from .lumfuncmcmc import LumFuncMCMC
        LFmod = LumFuncMCMC(z, fix_comp=args.fix_comp, Flim_lims=args.Flim_lims, alpha_lims=args.alpha_lims,
                            diff_rand=not args.same_rand, **common)
        print("Initialized LumFuncMCMC class")
    tag = _tag(args)
    imgtype = args.output_dict['image format']
    # result cache: an existing fitposterior file is reloaded and only re-summarised (reference run_lumfuncmcmc.py:259-270)
    fn = '%s/fitposterior_%s.dat' % (outdir, tag)
    if op.isfile(fn):
        LFmod.samples = Table.read(fn, format='ascii').as_array()
        LFmod.triangle_plot('%s/triangle_%s' % (outdir, tag), imgtype=imgtype)
        return LFmod
    names = LFmod.get_param_names()
    percentiles = args.param_percentiles
    labels = ['Line'] + [name + '_%02d' % per for name in names for per in percentiles]
    formats = {label: '%0.3f' for label in labels}
    formats['Line'] = '%s'
    print('Labels:', labels)
    LFmod.table = Table(names=labels, dtype=['S10'] + ['f8'] * (len(labels) - 1))
    print("Finished making names and labels for LF table and about to start fitting the model!")
    LFmod.fit_model()
    print("Finished fitting model and about to create outputs")
    if args.output_dict['triangle plot']:
        LFmod.triangle_plot('%s/triangle_%s' % (outdir, tag), imgtype=imgtype)
        print("Finished making Triangle Plot with Best-fit LF (and V_eff-method-based data)")
    else:
        LFmod.set_median_fit()
        print("Finished setting median fit and V_eff parameters")
    names.append('Ln Prob')
    fmt = 'ascii.fixed_width_two_line'
    if args.output_dict['fitposterior']:
        Table(LFmod.samples, names=names).write(fn, overwrite=True, format=fmt)
        print("Finished writing fitposterior file")
    if args.output_dict['bestfitLF']:
        if evolving:
            T = Table([np.repeat(LFmod.Lout[None], len(LFmod.zout), axis=0).ravel(),
                       np.repeat(LFmod.zout, len(LFmod.Lout)), LFmod.medianLF.ravel()],
                      names=['Luminosity_cols', 'Redshift_rows', 'MedianLFMatrix'])
        else:
            T = Table([LFmod.lum, LFmod.lum_e, LFmod.medianLF], names=['Luminosity', 'Luminosity_Err', 'MedianLF'])
        T.write('%s/bestfitLF_%s.dat' % (outdir, tag), overwrite=True, format=fmt)
        print("Finished writing bestfitLF file")
    if args.output_dict['VeffLF']:
        Table([LFmod.Lavg, LFmod.lfbinorig, np.sqrt(LFmod.var)], names=['Luminosity', 'BinLF', 'BinLFErr']).write(
            '%s/VeffLF_%s.dat' % (outdir, tag), overwrite=True, format=fmt)
        print("Finished writing VeffLF file")
    LFmod.table.add_row([args.line_name] + [0.] * (len(labels) - 1))
    LFmod.add_fitinfo_to_table(percentiles)
    print(LFmod.table)
    if args.output_dict['parameters']:
        LFmod.table.write('%s/%s' % (outdir, args.output_filename), format=fmt, formats=formats, overwrite=True)
        print("Finished writing LF main table")
    if args.output_dict['settings']:
        del args.log
        with open('%s/%s.args' % (outdir, args.output_filename), 'w') as fh:
            fh.write(str(vars(args)))
        print("Finished writing settings to file")
    return LFmod
