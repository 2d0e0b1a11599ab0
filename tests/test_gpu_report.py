"""-m gpu: report aggregation on the device tables (sequali_b200.report, csrc/report.cu) against
the reference's own report_modules.py.  The reference side runs in a subprocess: its unchanged package files
(staged by oracle/build_ref.sh in the git-ignored oracle/_ref/pkg_src) around its compiled extension, with a
stand-in for the absent `pygal` (nothing is plotted).  Floats are compared by bit pattern."""
import io
import json
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

from sequali_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

PKG_SRC = os.path.join(H.ROOT, "oracle", "_ref", "pkg_src")
REF_SO = os.path.join(H.ROOT, "oracle", "_ref", "sequali", "_qc.abi3.so")
SHIM = os.path.join(H.ROOT, "oracle", "_ref", "tests", "shim")
if not (os.path.exists(os.path.join(PKG_SRC, "report_modules.py")) and os.path.exists(REF_SO)):  # pragma: no cover
    pytest.skip("oracle/build_ref.sh has not staged report_modules.py", allow_module_level=True)

REFERENCE_SIDE = r'''
import dataclasses, json, sys
from sequali import report_modules as rm
from sequali._qc import FastqParser, BamParser, QCMetrics, NanoStats

path, bam = sys.argv[1], sys.argv[2] == "bam"
metrics, nano = QCMetrics(), NanoStats()
with open(path, "rb") as f:
    for arr in (BamParser(f) if bam else FastqParser(f)):
        metrics.add_record_array(arr)
        nano.add_record_array(arr)

def enc(x):
    if isinstance(x, float):
        return {"f": x.hex()}
    if isinstance(x, dict):
        return {"d": [[enc(k), enc(v)] for k, v in x.items()]}
    if isinstance(x, (list, tuple)):
        return [enc(v) for v in x]
    return x

max_length = metrics.max_length
ranges = list(rm.logarithmic_ranges(max_length)) if max_length > 500 else list(rm.equidistant_ranges(max_length, 200))
modules = rm.qc_metrics_modules(metrics, ranges)
base, phred = metrics.base_count_table(), metrics.phred_count_table()
out = {
    "ranges": ranges,
    "aggregated_base_matrix": list(rm.aggregate_count_matrix(base, ranges, 5)),
    "aggregated_phred_matrix": list(rm.aggregate_count_matrix(phred, ranges, 12)),
    "summary": enc(dataclasses.asdict(modules[0])),
    "sequence_length_distribution": enc(dataclasses.asdict(modules[1])),
    "nanostats": enc(dataclasses.asdict(rm.NanoStatsReport.from_nanostats(nano))),
}
print(json.dumps(out))
'''


def enc(x):
    if isinstance(x, float):
        return {"f": x.hex()}
    if isinstance(x, dict):
        return {"d": [[enc(k), enc(v)] for k, v in x.items()]}
    if isinstance(x, (list, tuple)):
        return [enc(v) for v in x]
    return x


@pytest.fixture(scope="module")
def ref_env(tmp_path_factory):
    root = tmp_path_factory.mktemp("refpkg")
    pkg = root / "sequali"
    shutil.copytree(PKG_SRC, pkg)
    for so in os.listdir(os.path.dirname(REF_SO)):
        if so.endswith(".so"):
            shutil.copy(os.path.join(os.path.dirname(REF_SO), so), pkg / so)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(root), SHIM])
    return env


@pytest.fixture(scope="module", params=["ctypes", "extension"])
def sq(request):
    if request.param == "extension":
        import sequali_b200.ext
        return sequali_b200.ext
    import sequali_b200
    return sequali_b200


def reference_report(ref_env, tmp_path, data, bam):
    path = tmp_path / "input"
    path.write_bytes(data)
    proc = subprocess.run([sys.executable, "-c", REFERENCE_SIDE, str(path), "bam" if bam else "fastq"], env=ref_env,
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stderr[-3000:]
    return json.loads(proc.stdout)


def ours(sq, data, bam):
    from sequali_b200 import report
    metrics, nano = sq.QCMetrics(), sq.NanoStats()
    parser = sq.BamParser(io.BytesIO(data)) if bam else sq.FastqParser(io.BytesIO(data))
    for arr in parser:
        metrics.add_record_array(arr)
        nano.add_record_array(arr)
    ranges = report.data_ranges_for(metrics.max_length)
    return ranges, report.qc_metrics_tables(metrics, ranges), report.nanostats_report(nano)


def check(sq, ref_env, tmp_path, data, bam=False):
    want = reference_report(ref_env, tmp_path, data, bam)
    ranges, qc, nano = ours(sq, data, bam)
    assert [list(r) for r in ranges] == want["ranges"]
    assert list(qc["aggregated_base_matrix"]) == want["aggregated_base_matrix"]
    assert list(qc["aggregated_phred_matrix"]) == want["aggregated_phred_matrix"]
    w = dict((k[0], k[1]) for k in want["summary"]["d"])
    for key, value in qc["summary"].items():
        assert enc(value) == w[key], key
    w = dict((k[0], k[1]) for k in want["sequence_length_distribution"]["d"])
    for key, value in qc["sequence_length_distribution"].items():
        assert enc(value) == w[key], (key, value, w[key])
    w = dict((k[0], k[1]) for k in want["nanostats"]["d"])
    assert set(w) == set(nano)
    for key, value in nano.items():
        assert enc(value) == w[key], key
    return qc, nano


def test_nanopore_fastq_headers(sq, ref_env, tmp_path):
    data = synth.nanopore_fastq(3000, mean_length=3000, max_length=60_000, seed=71)
    qc, nano = check(sq, ref_env, tmp_path, data)
    assert len(nano["x_labels"]) > 100 and sum(nano["time_reads"]) == 3000 and len(nano["per_channel_bases"]) > 500
    assert qc["sequence_length_distribution"]["n50"] > qc["sequence_length_distribution"]["q50"] > 0


def test_nanopore_ubam_tags(sq, ref_env, tmp_path):
    data = synth.nanopore_ubam(2500, mean_length=2500, max_length=80_000, seed=72)
    qc, nano = check(sq, ref_env, tmp_path, data, bam=True)
    assert sum(nano["translocation_speed"]) > 2000 and nano["reads_with_parent"]


def test_illumina_equidistant_ranges_and_skipped_nanostats(sq, ref_env, tmp_path):
    data = synth.illumina_fastq(20_000, 151, seed=73, n_tiles=8)
    qc, nano = check(sq, ref_env, tmp_path, data)
    assert nano["skipped_reason"] and qc["summary"]["minimum_length"] == 151


def ragged_fastq(rng, n, lengths):
    out = io.BytesIO()
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for i in range(n):
        ln = int(lengths[i])
        seq = letters[rng.integers(0, 5, size=ln)].tobytes()
        qual = (rng.integers(0, 60, size=ln).astype(np.uint8) + 33).tobytes()
        out.write(b"@r%d\n" % i + seq + b"\n+\n" + qual + b"\n")
    return out.getvalue()


@pytest.mark.parametrize("case", ["geometric", "with_empty_reads", "one_read", "two_lengths", "long_tail"])
def test_length_distribution_walk(sq, ref_env, tmp_path, case):
    rng = np.random.default_rng(74)
    lengths = {"geometric": rng.geometric(0.02, size=4000),
               "with_empty_reads": np.concatenate([np.zeros(300, dtype=int), rng.integers(0, 40, size=700)]),
               "one_read": np.array([77]),
               "two_lengths": np.array([10] * 99 + [700]),
               "long_tail": np.concatenate([rng.integers(1, 30, size=500), np.array([5000, 9000])])}[case]
    check(sq, ref_env, tmp_path, ragged_fastq(rng, len(lengths), lengths))


def test_no_reads(sq, ref_env, tmp_path):
    check(sq, ref_env, tmp_path, b"")


def test_odd_durations_and_channels(sq, ref_env, tmp_path):
    """Negative channel ids sort in front; a zero duration is skipped, a negative one lands at the far end of the
    translocation table like a negative Python index does."""
    rng = np.random.default_rng(75)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = io.BytesIO()
    out.write(synth.bam_header())
    for i in range(400):
        ln = int(rng.integers(1, 3000))
        dur = [0.0, -4.0, 0.004, 2.25, 1e9][i % 5]
        tags = b"chi" + struct.pack("<i", int(rng.integers(-5, 6))) + b"stZ2024-03-0%dT0%d:00:00Z\0" % (i % 9 + 1, i % 7) + \
            b"duf" + struct.pack("<f", dur)
        out.write(synth.bam_record(b"r%d" % i, letters[rng.integers(0, 4, ln)], rng.integers(33, 80, ln).astype(np.uint8), tags))
    qc, nano = check(sq, ref_env, tmp_path, out.getvalue(), bam=True)
    assert list(nano["per_channel_bases"])[0] < 0
