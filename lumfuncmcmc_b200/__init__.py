"""lumfuncmcmc_b200 -- B200-native likelihood engine behind LumFuncMCMC's Python surface."""
__version__ = "0.1.0"
