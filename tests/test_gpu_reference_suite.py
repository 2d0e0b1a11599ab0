"""-m gpu: the reference's OWN hot-path test files, unmodified, run against the B200 build.

SURVEY.md 8(c): the ten test files below (615 cases) pass against the compiled reference and must
pass against the drop-in module.  `oracle/build_ref.sh` stages them -- with `tests/data` and a stdlib
`xopen` stand-in -- in the git-ignored `oracle/_ref/tests/` (they travel to the GPU box with the
snapshot, never into the repo's history).  Here they run in a subprocess whose `sequali` package is
an alias of the B200 build: `sequali._qc` is the CPython extension `sequali_b200/_qc_ext` when it
is built (SEQUALI_B200_EXT=1, the default when the .so is present), else the ctypes mirror.
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref", "tests")

ALIAS = '''\
# alias package written by tests/test_gpu_reference_suite.py
import sys
import {impl} as _impl
from {impl} import *  # noqa: F401,F403
from {impl} import _qc  # noqa: F401
sys.modules[__name__ + "._qc"] = _qc
__version__ = "b200"
'''


PKG_SRC = os.path.join(ROOT, "oracle", "_ref", "pkg_src")
EXT_SO = os.path.join(ROOT, "sequali_b200", "ext", "_qc.so")


def run_suite(tmp_path, impl):
    pkg = tmp_path / "alias" / "sequali"
    if impl == "sequali":
        # the reference's own package files, unchanged, around the B200 build's extension module
        import shutil
        shutil.copytree(PKG_SRC, pkg)
        shutil.copy(EXT_SO, pkg / "_qc.so")
    else:
        pkg.mkdir(parents=True)
        (pkg / "__init__.py").write_text(ALIAS.format(impl=impl))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path / "alias"), os.path.join(STAGE, "shim"), ROOT])
    proc = subprocess.run([sys.executable, "-m", "pytest", "tests", "-q", "-p", "no:cacheprovider",
                           "-o", "addopts=", "--tb=line", "-W", "ignore::DeprecationWarning"],
                          cwd=STAGE, env=env, capture_output=True, text=True, timeout=3000)
    out = proc.stdout + proc.stderr
    m = re.search(r"(\d+) passed", out)
    passed = int(m.group(1)) if m else 0
    failed = re.findall(r"^(?:FAILED|ERROR) (\S+)", out, flags=re.M)
    return proc.returncode, passed, failed, out


@pytest.mark.skipif(not os.path.isdir(os.path.join(STAGE, "tests")),
                    reason="oracle/build_ref.sh has not staged the reference's tests")
@pytest.mark.parametrize("impl", ["sequali_b200", "sequali_b200.ext", "sequali"])
def test_reference_hot_path_tests_pass(tmp_path, impl):
    """impl = the ctypes mirror / the extension under an alias package / the reference's unchanged
    __init__.py + util.py + adapters.py with the extension as `sequali._qc`."""
    if impl == "sequali" and not (os.path.isdir(PKG_SRC) and os.path.exists(EXT_SO)):
        pytest.skip("package shell or extension not staged")
    rc, passed, failed, out = run_suite(tmp_path, impl)
    assert rc == 0 and not failed, f"{passed} passed, {len(failed)} failed:\n" + "\n".join(failed[:40]) + \
        "\n" + out[-3000:]
    assert passed >= 615, out[-2000:]
