"""Every Python file of the repository byte-compiles, and every CUDA / C++ tool names files that exist (the tools are run by
hand on GPU boxes, so nothing else would notice a stale import or path)."""
import glob
import os
import py_compile
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_python_files_compile(tmp_path):
    files = [f for pat in ('*.py', 'lumfuncmcmc_b200/*.py', 'oracle/*.py', 'tools/*.py', 'tools/math/*.py', 'tests/*.py')
             for f in glob.glob(os.path.join(ROOT, pat))]
    assert len(files) > 40
    for f in files:
        py_compile.compile(f, cfile=str(tmp_path / (os.path.basename(f) + 'c')), doraise=True)


def test_profile_script_and_docs_reference_existing_files():
    """Paths quoted in the evidence script and in profiles/README.md exist in the tree."""
    script = open(os.path.join(ROOT, 'tools', 'r2_profile.sh')).read()
    for tool in re.findall(r'python (tools/[\w/]+\.py)', script):
        assert os.path.isfile(os.path.join(ROOT, tool)), tool
    readme = open(os.path.join(ROOT, 'profiles', 'README.md')).read()
    round2 = readme[readme.index('## Round 2'):readme.index('## Round 1')]
    missing = [name for name in set(re.findall(r'`(r02_[\w.]+\.(?:json|txt|csv|md|log))`', round2))
               if not os.path.isfile(os.path.join(ROOT, 'profiles', name))]
    assert not missing, missing
