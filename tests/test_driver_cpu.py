"""CPU: driver-side host logic -- flag/config fallback, catalogue reader, table formats (no GPU needed)."""
import numpy as np

from lumfuncmcmc_b200 import configLF, synth
from lumfuncmcmc_b200.driver import parse_args, read_input_file
from lumfuncmcmc_b200.tableio import Table


def _write_catalogue(path, cat, line='OIII'):
    with open(path, 'w') as fh:
        fh.write('Field z ID %s_flux %s_flux_e\n' % (line, line))
        i = 0
        for k, name in enumerate(cat['field_names']):
            for z, f, fe in zip(cat['z'][k], cat['flux'][k], cat['flux_e'][k]):
                fh.write('%s %.17g %d %.17g %.17g\n' % (name, z, i, f, fe))
                i += 1


def test_flags_fall_back_to_config():
    a = parse_args(['-f', 'x.dat', '-nw', '64'])
    assert a.nwalkers == 64 and a.nsteps == configLF.nsteps and a.Flim == configLF.Flim
    assert a.alpha == configLF.alpha and a.min_comp_frac == configLF.min_comp_frac
    assert a.fix_comp == 0 and a.fix_sch_al == 0 and a.same_rand == 0          # count flags stay 0 unless given
    assert a.Flim_lims == configLF.Flim_lims and a.line_plot_name.startswith('[OIII]')
    b = parse_args(['-f', 'x.dat', '-fc', '-fsa', '-ln', 'Ha'])
    assert b.fix_comp == 1 and b.fix_sch_al == 1 and 'alpha' in b.line_plot_name
    z = parse_args(['-f', 'x.dat'], evolving=True)
    assert (z.z1, z.z2, z.z3) == (1.20, 1.53, 1.86) and not hasattr(z, 'fix_comp')
    assert parse_args(['-f', 'OIIIGautamWillNoAGN.dat'], evolving=True).z2 == 1.76
    assert parse_args(['-f', 'x.dat', '-ln', 'Ha'], evolving=True).z3 == 1.54


def test_catalogue_reader_groups_by_field(tmp_path):
    cat = synth.make_catalogue(300, seed=3, nfields=3)
    path = str(tmp_path / 'cat.dat')
    _write_catalogue(path, cat)
    args = parse_args(['-f', path])
    z, flux, flux_e, lum, lum_e, names, field_ind = read_input_file(args)
    assert list(names) == ['F0', 'F1', 'F2'] and lum is None
    assert np.array_equal(field_ind, cat['field_ind'])
    for k in range(3):
        assert np.array_equal(z[k], cat['z'][k]) and np.array_equal(flux[k], cat['flux'][k])
        assert np.array_equal(flux_e[k], cat['flux_e'][k])


def test_fixed_width_two_line_round_trip(tmp_path):
    samples = np.random.default_rng(1).normal(size=(7, 3))
    path = str(tmp_path / 'post.dat')
    Table(samples, names=['$\\log L_*$', '$\\alpha$', 'Ln Prob']).write(path, format='ascii.fixed_width_two_line')
    lines = open(path).read().splitlines()
    assert set(lines[1]) <= set('- ') and len(lines) == 9
    back = Table.read(path, format='ascii')
    assert back.colnames[-1] == 'Ln Prob'
    assert np.array_equal(back.as_array(), samples)                  # repr round-trips float64 exactly
    t = Table(names=['Line', 'a_16', 'a_50'], dtype=['S10', 'f8', 'f8'])
    t.add_row(['OIII', 0., 0.])
    t[-1][1] = 1.23456
    t[-1][2] = 7.0
    assert len(t[0]) == 3
    t.write(str(tmp_path / 'par.dat'), formats={'Line': '%s', 'a_16': '%0.3f', 'a_50': '%0.3f'})
    assert '1.235' in open(str(tmp_path / 'par.dat')).read()
