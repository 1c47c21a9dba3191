"""Accuracy of the streaming log / exp routines of the 1/V_eff kernel, measured on a host replica that performs the
same operations in the same order (tools/math/stream_accuracy.cpp).  CPU only: needs g++."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None, reason="needs g++")
def test_streaming_log_and_exp_are_libm_grade(tmp_path):
    exe = str(tmp_path / 'stream_accuracy')
    subprocess.run(['g++', '-O2', '-o', exe, os.path.join(ROOT, 'tools', 'math', 'stream_accuracy.cpp')], check=True)
    out = subprocess.run([exe, '2000000'], check=True, capture_output=True, text=True).stdout
    # "log: max abs err A, max err / max(|ln|, 1) B; exp: max rel err C"
    nums = [float(tok.rstrip(',;')) for tok in out.replace(';', ' ').split() if tok[0].isdigit() and 'e-' in tok]
    assert len(nums) == 3, out
    log_abs, log_rel, exp_rel = nums
    assert log_rel < 3.0e-16 and exp_rel < 3.0e-16 and log_abs < 1.0e-14, out


@pytest.mark.skipif(shutil.which('g++') is None, reason="needs g++")
def test_fast_term_of_the_free_model_stays_inside_its_error_budget(tmp_path):
    """One walker x source term of k_main<false, FREE> (math v6) on its host replica: relative deviation from long double
    arithmetic below 3e-12 for MUFU seeds anywhere inside their measured bounds (tolerance on lnprob: 1e-10 relative)."""
    exe = str(tmp_path / 'term_accuracy')
    subprocess.run(['g++', '-O2', '-o', exe, os.path.join(ROOT, 'tools', 'math', 'term_accuracy.cpp')], check=True)
    out = subprocess.run([exe, '2000000'], check=True, capture_output=True, text=True).stdout
    rel = float(out.split('max err / max(|t|, 1)')[1].split(',')[0])
    assert rel < 3.0e-12, out


def test_replica_and_kernel_header_use_the_same_constants():
    """The host replica restates lf_math.cuh's constants by hand; a coefficient changed in one place only would make the
    error budget above describe a different polynomial than the one the kernel runs."""
    import re
    hdr = open(os.path.join(ROOT, 'lumfuncmcmc_b200', 'csrc', 'lf_math.cuh')).read()
    rep = open(os.path.join(ROOT, 'tools', 'math', 'term_accuracy.cpp')).read()

    def const(text, name):
        m = re.search(r'\b%s\s*=\s*([-+0-9a-fA-FxXpP.]+)' % name, text)
        assert m, name
        tok = m.group(1)
        return float.fromhex(tok) if tok.lower().startswith(('0x', '-0x')) else float(tok)

    for name in ('MAGIC44', 'LOG1P_C0', 'LOG1P_C2', 'LOG1P_C3_HI', 'EXP2_C0', 'EXP2N_C1', 'EXP2N_C2', 'EXP2N_C3_HI'):
        assert const(hdr, name) == const(rep, name), name
    # the cubic coefficients are DFMA immediates: low 32 bits of the double must be zero
    import struct
    for name in ('LOG1P_C3_HI', 'EXP2N_C3_HI', 'MAGIC44'):
        assert struct.unpack('<Q', struct.pack('<d', const(hdr, name)))[0] & 0xffffffff == 0, name
