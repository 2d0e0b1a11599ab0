"""-m gpu: `_seqident.sequence_identity` of the B200 build (k_seqident) against the compiled, unmodified reference
module (oracle/_ref/sequali/_seqident, built by oracle/build_ref.sh from _seqidentmodule.c): SURVEY.md 8(f)4."""
import importlib
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _reference():
    if H.import_reference() is None:
        return None
    try:
        return importlib.import_module("sequali._seqident")
    except ImportError:
        return None


REF = _reference()
if REF is None:  # pragma: no cover
    pytest.skip("oracle/_ref/sequali/_seqident is not built", allow_module_level=True)


@pytest.fixture(scope="module", params=["ctypes", "extension"])
def si(request):
    if request.param == "extension":
        from sequali_b200.ext import _seqident
        return _seqident
    from sequali_b200 import _seqident
    return _seqident


def same(a, b):
    return a == b or (math.isnan(a) and math.isnan(b))


def random_pairs(rng, n):
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    pairs = []
    for k in range(n):
        t_len = int(rng.integers(0, 300)) if k % 50 else int(rng.integers(0, 4))
        target = letters[rng.choice(5, size=t_len, p=[.24, .24, .24, .24, .04])]
        q_len = int(rng.integers(0, 32))
        if k % 3 and t_len > q_len:  # a mutated piece of the target: substitutions, one deletion, one insertion
            at = int(rng.integers(0, t_len - q_len + 1))
            q = target[at:at + q_len].copy()
            for _ in range(int(rng.integers(0, 4))):
                if len(q):
                    q[int(rng.integers(0, len(q)))] = letters[int(rng.integers(0, 4))]
            if len(q) > 4 and rng.random() < .4:
                q = np.delete(q, int(rng.integers(1, len(q) - 1)))
            if 4 < len(q) < 31 and rng.random() < .4:
                q = np.insert(q, int(rng.integers(1, len(q) - 1)), letters[int(rng.integers(0, 4))])
        else:
            q = letters[rng.integers(0, 4, size=q_len)]
        pairs.append((target.tobytes().decode(), q.tobytes().decode()))
    return pairs


def test_known_alignments(si):
    # answers worked out by hand from the recurrence (an inserted target letter costs one match, a skipped
    # query letter none), each also put to the reference
    cases = [("GGGACGTGGG", "ACGT", 1.0), ("ACGT", "ACGT", 1.0), ("ACTGT", "ACGT", 3 / 4), ("ACGT", "ACTGT", 4 / 5),
             ("ACGT", "AGGT", 3 / 4), ("", "ACGT", 0.0), ("TTTT", "ACG", 0.0), ("A", "A", 1.0)]
    for target, query, want in cases:
        assert REF.sequence_identity(target, query) == want, (target, query)
        assert si.sequence_identity(target, query) == want, (target, query)
    assert math.isnan(si.sequence_identity("ACGT", "")) and math.isnan(REF.sequence_identity("ACGT", ""))


def test_random_pairs_default_scores(si):
    pairs = random_pairs(np.random.default_rng(61), 3000)
    want = [REF.sequence_identity(t, q) for t, q in pairs]
    got = si.sequence_identities(pairs)
    bad = [(p, g, w) for p, g, w in zip(pairs, got, want) if not same(g, w)]
    assert not bad, bad[:5]
    for t, q in pairs[:40]:
        assert same(si.sequence_identity(t, q), REF.sequence_identity(t, q))


@pytest.mark.parametrize("scores", [(2, -1, -2, -1), (1, -3, -2, -2), (3, -2, -1, -4), (1, 0, -1, -1), (2, -2, -3, -3)])
def test_random_pairs_other_scores(si, scores):
    pairs = random_pairs(np.random.default_rng(62), 800)
    kw = dict(zip(("match_score", "mismatch_penalty", "deletion_penalty", "insertion_penalty"), scores))
    want = [REF.sequence_identity(t, q, **kw) for t, q in pairs]
    got = si.sequence_identities(pairs, **kw)
    bad = [(p, g, w) for p, g, w in zip(pairs, got, want) if not same(g, w)]
    assert not bad, bad[:5]
    assert same(si.sequence_identity(*pairs[7], *scores), want[7])  # positional scores


def test_long_target(si):
    rng = np.random.default_rng(63)
    target = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 200_000)].tobytes().decode()
    query = target[150_000:150_031]
    assert si.sequence_identity(target, query) == REF.sequence_identity(target, query) == 1.0
    query = "ACGTTGCAAGGCTTAACCGGTTAACGTACGT"
    assert si.sequence_identity(target, query) == REF.sequence_identity(target, query)


def test_errors_are_the_references(si):
    for args in [("ACGT", "A" * 32), ("ACé", "AC"), ("ACGT", "é"), (b"ACGT", "A"), ("ACGT", 5), ("ACGT",)]:
        with pytest.raises(Exception) as want:
            REF.sequence_identity(*args)
        with pytest.raises(Exception) as got:
            si.sequence_identity(*args)
        assert type(got.value) is type(want.value), args
        if len(args) == 2:
            assert str(got.value) == str(want.value), args


def test_reference_test_file_and_identify_sequence_under_the_unchanged_package(tmp_path):
    """The reference's test_sequence_identification.py and its identify_sequence_builtin (k-mer index over the
    contaminant library, then sequence_identity per candidate), with the reference's unchanged package files around
    the B200 build's _qc.so / _seqident.so."""
    import shutil
    pkg_src = os.path.join(H.ROOT, "oracle", "_ref", "pkg_src")
    stage = os.path.join(H.ROOT, "oracle", "_ref", "tests_seqident")
    ext = os.path.join(H.ROOT, "sequali_b200", "ext")
    if not (os.path.isdir(stage) and os.path.exists(os.path.join(pkg_src, "sequence_identification.py"))):
        pytest.skip("oracle/build_ref.sh has not staged the package shell")
    pkg = tmp_path / "alias" / "sequali"
    shutil.copytree(pkg_src, pkg)
    shutil.copy(os.path.join(ext, "_qc.so"), pkg / "_qc.so")
    shutil.copy(os.path.join(ext, "_seqident.so"), pkg / "_seqident.so")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path / "alias"), os.path.join(H.ROOT, "oracle", "_ref", "tests", "shim"), H.ROOT])
    proc = subprocess.run([sys.executable, "-m", "pytest", "tests", "-q", "-p", "no:cacheprovider", "-o", "addopts=",
                           "--tb=short"], cwd=stage, env=env, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and " passed" in proc.stdout and "failed" not in proc.stdout, proc.stdout[-3000:] + proc.stderr[-2000:]
    # identify_sequence_builtin: same answers as the reference's own module for adapter-like and random sequences
    probe = ("import json, sys\n"
             "from sequali.sequence_identification import identify_sequence_builtin\n"
             "seqs = ['AGATCGGAAGAGCACACGTCTGAACTCCAGT', 'AGATCGGAAGAGCGTCGTGTAGGGAAAGAGT', 'CTGTCTCTTATACACATCT',\n"
             "        'TTTTTTTTTTTTTTTTTTTTT', 'GATCGGAAGAGCACACGTCTG', 'ACGTTGCAAGGCTTAACCGGT', 'AATGATACGGCGACCACCGAG']\n"
             "print(json.dumps([identify_sequence_builtin(s) for s in seqs]))\n")
    got = subprocess.run([sys.executable, "-c", probe], env=env, capture_output=True, text=True, timeout=900)
    assert got.returncode == 0, got.stderr[-2000:]
    ref_pkg = tmp_path / "ref" / "sequali"
    shutil.copytree(pkg_src, ref_pkg)
    for so in ("_qc.abi3.so", "_seqident.abi3.so"):
        shutil.copy(os.path.join(H.ROOT, "oracle", "_ref", "sequali", so), ref_pkg / so)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path / "ref"), os.path.join(H.ROOT, "oracle", "_ref", "tests", "shim")])
    want = subprocess.run([sys.executable, "-c", probe], env=env, capture_output=True, text=True, timeout=900)
    assert want.returncode == 0, want.stderr[-2000:]
    assert got.stdout == want.stdout and "llumina" in want.stdout
