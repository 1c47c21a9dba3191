#!/usr/bin/env python
"""Single-redshift Schechter + Fleming-completeness fit, same command line and outputs as the reference's
run_lumfuncmcmc.py (reference run_lumfuncmcmc.py:230-330), likelihood on the B200 engine.

    python run_lumfuncmcmc.py -f catalogue.dat -o fit.dat [-nw 100 -ns 1000 -fc -fsa -mcf 0.0 ...] [--device 0]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lumfuncmcmc_b200.driver import parse_args as _parse, read_input_file, run   # noqa: E402,F401


def parse_args(argv=None):
    return _parse(argv, evolving=False)


def main(argv=None):
    return run(argv, evolving=False, script_name='run_lumfuncmcmc.py')


if __name__ == '__main__':
    main()
