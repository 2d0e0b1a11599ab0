"""CPU: host logic of the rows beside the hot path -- the report's position ranges against the
reference's own functions, the argument handling of `_seqident`, and that none of it answers without a GPU."""
import os
import subprocess
import sys

import pytest

from tests import helpers as H

PKG_SRC = os.path.join(H.ROOT, "oracle", "_ref", "pkg_src")
SHIM = os.path.join(H.ROOT, "oracle", "_ref", "tests", "shim")


def test_report_ranges_equal_the_references(tmp_path):
    """equidistant_ranges / logarithmic_ranges / stringify_ranges of sequali_b200.report against report_modules.py
    (run in a subprocess around the compiled reference extension; pygal stubbed)."""
    ref_so = os.path.join(H.ROOT, "oracle", "_ref", "sequali", "_qc.abi3.so")
    if not (os.path.exists(os.path.join(PKG_SRC, "report_modules.py")) and os.path.exists(ref_so)):
        pytest.skip("oracle/build_ref.sh has not staged report_modules.py")
    import shutil
    pkg = tmp_path / "sequali"
    shutil.copytree(PKG_SRC, pkg)
    for so in os.listdir(os.path.dirname(ref_so)):
        if so.endswith(".so"):
            shutil.copy(os.path.join(os.path.dirname(ref_so), so), pkg / so)
    code = ("import json\nfrom sequali import report_modules as rm\n"
            "lengths = [0, 1, 2, 5, 150, 151, 199, 200, 201, 250, 499, 500, 501, 1000, 20000, 1000000, 250000000]\n"
            "out = {}\n"
            "for n in lengths:\n"
            "    r = list(rm.logarithmic_ranges(n)) if n > 500 else list(rm.equidistant_ranges(n, 200))\n"
            "    out[str(n)] = [r, rm.stringify_ranges(r)]\n"
            "print(json.dumps(out))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), SHIM]))
    proc = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    import json
    want = json.loads(proc.stdout)
    from sequali_b200 import report
    for n, (ranges, labels) in want.items():
        got = report.data_ranges_for(int(n))
        assert [list(r) for r in got] == ranges, n
        assert report.stringify_ranges(got) == labels, n


def test_seqident_checks_its_arguments_before_it_needs_a_device():
    from sequali_b200 import _seqident as mirror
    from sequali_b200.ext import _seqident as ext
    for mod in (mirror, ext):
        with pytest.raises(ValueError, match="Only query with lengths less than 32 are supported. Got 32"):
            mod.sequence_identity("ACGT", "A" * 32)
        with pytest.raises(ValueError, match="Only ascii strings are allowed"):
            mod.sequence_identity("ACGé", "ACG")
        with pytest.raises(TypeError, match="argument 1 must be str, not bytes"):
            mod.sequence_identity(b"ACGT", "A")
        assert mod.sequence_identities([]) == []
    assert mirror._int8(-1) == -1 and mirror._int8(127) == 127 and mirror._int8(128) == -128 and mirror._int8(300) == 44


@pytest.mark.skipif(H.gpu_available() if hasattr(H, "gpu_available") else False, reason="a GPU is present")
def test_next_rows_fail_loudly_without_a_gpu():
    """No CPU fallback anywhere: the calls that would need the device raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sequali_b200.ext import _seqident as ext
    with pytest.raises(RuntimeError, match="no CUDA device|needs a GPU"):
        ext.sequence_identity("ACGT", "ACG")
    from sequali_b200 import _seqident as mirror
    with pytest.raises(Exception):
        mirror.sequence_identity("ACGT", "ACG")
