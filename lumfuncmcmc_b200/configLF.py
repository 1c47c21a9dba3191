"""Run defaults, same attribute names as the reference's configLF.py (reference configLF.py:1-42).

The drivers fall back to these whenever a command-line option is left unset (reference
run_lumfuncmcmc.py:120-127), so the names are part of the public surface.
"""
# sampler
nwalkers, nsteps = 100, 1000
# 1/V_eff binned luminosity function
nbins, nboot = 50, 100
# modified-Fleming completeness: per-field 50% flux (x1e-17 erg/cm^2/s), slope, prior boxes
Flim = [2.72, 3.61, 2.55, 3.31, 3.30]
Flim_lims = [1.0, 6.0]
alpha = 4.56
alpha_lims = [1.0, 7.0]
fcmin = 0.1
min_comp_frac = 0.0
# line
line_name = "OIII"
line_plot_name = r'[OIII] $\lambda 5007$'
# effective areas: arcmin^2 -> arcsec^2 with a usable fraction
Omega_0_sqarcmin = [121.9, 122.2, 116.0, 147.3, 118.7]
frac_use, conv_minsec = 0.85, 3600
Omega_0 = [val * frac_use * conv_minsec for val in Omega_0_sqarcmin]
# Schechter parameters and prior boxes
sch_al, sch_al_lims = -1.49, [-3.0, 1.0]
Lstar, Lstar_lims = 42.5, [40.0, 45.0]
phistar, phistar_lims = -2.0, [-8.0, 5.0]
# luminosity range of the likelihood integral
Lc, Lh = 40.0, 46.0
# reported percentiles and output switches
param_percentiles = [5, 16, 50, 84, 95]
output_dict = {'parameters': True, 'settings': True, 'fitposterior': True, 'bestfitLF': True,
               'VeffLF': True, 'triangle plot': True, 'image format': 'png'}
