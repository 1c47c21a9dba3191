#!/usr/bin/env python
"""Redshift-evolving luminosity-function fit, same command line and outputs as the reference's run_lumfuncmcmc_z.py
(reference run_lumfuncmcmc_z.py:203-303), likelihood on the B200 engine.

    python run_lumfuncmcmc_z.py -f catalogue.dat -o fitz.dat [-nw 100 -ns 1000 -fsa ...] [--device 0]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lumfuncmcmc_b200.driver import parse_args as _parse, read_input_file, run   # noqa: E402,F401


def parse_args(argv=None):
    return _parse(argv, evolving=True)


def main(argv=None):
    return run(argv, evolving=True, script_name='run_lumfuncmcmc_z.py')


if __name__ == '__main__':
    main()
