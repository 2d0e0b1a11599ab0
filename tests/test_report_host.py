"""CPU: the host half of sequali_b200.report (what it does with the numbers the device kernels return) and the
telescoping formulas of csrc/report.cu, restated in numpy, against the reference's report_modules.py.

`FakeMetrics.aggregate` / `FakeNano.report_tables` below are the numpy restatement of `sq_qc_aggregate` /
`sq_nanostats_report` (test infrastructure: the product calls the device); their inputs are the tables of the
compiled, unmodified reference collectors.  The reference's own report code runs in a subprocess (its unchanged
package files around its extension, `pygal` stubbed), as in tests/test_gpu_report.py."""
import array
import io
import json
import math
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from sequali_b200 import report, synth
from tests import helpers as H
from tests.test_gpu_report import REFERENCE_SIDE, enc

REF = H.import_reference()
PKG_SRC = os.path.join(H.ROOT, "oracle", "_ref", "pkg_src")
SHIM = os.path.join(H.ROOT, "oracle", "_ref", "tests", "shim")
if REF is None or not os.path.exists(os.path.join(PKG_SRC, "report_modules.py")):  # pragma: no cover
    pytest.skip("oracle/_ref (reference extension + staged package files) is not built", allow_module_level=True)


class FakeMetrics:
    """QCMetrics of the reference + aggregate() as csrc/report.cu computes it."""

    def __init__(self, qc):
        self.qc = qc
        self.number_of_reads, self.max_length = qc.number_of_reads, qc.max_length
        self.base = np.frombuffer(qc.base_count_table(), dtype=np.uint64).reshape(-1, 5)
        self.phred = np.frombuffer(qc.phred_count_table(), dtype=np.uint64).reshape(-1, 12)

    def phred_scores(self):
        return self.qc.phred_scores()

    def aggregate(self, data_ranges, count_thresholds=()):
        ml = self.max_length
        r = np.concatenate([self.base.sum(axis=1), [0]]).astype(object)          # r[i] = reads longer than i
        base, phred, lengths = [], [], []
        for start, stop in data_ranges:
            start, stop = min(start, ml), min(stop, ml)
            base.extend(int(v) for v in self.base[start:stop].sum(axis=0))
            phred.extend(int(v) for v in self.phred[start:stop].sum(axis=0))
            lengths.append(int(r[start] - r[stop]) if start < stop else 0)
        total = int(sum(r))
        prefix = [0]
        for v in r[:ml]:
            prefix.append(prefix[-1] + int(v))

        def first(pred):
            return next((length for length in range(ml + 1) if pred(length)), None)

        def bases_upto(length):
            return prefix[length] - length * int(r[length])

        n50 = first(lambda length: bases_upto(length) >= total // 2)
        n90 = first(lambda length: bases_upto(length) >= int(total * 0.1))
        minimum = first(lambda i: i < ml and int(r[i]) < self.number_of_reads)
        thr = [first(lambda length: int(r[0]) - int(r[length]) > t) or 0 for t in count_thresholds]
        return {"base_matrix": array.array("Q", base), "phred_matrix": array.array("Q", phred), "length_counts": lengths,
                "total_bases": total, "minimum_length": ml if minimum is None else minimum, "n50": n50, "n90": n90,
                "threshold_lengths": thr}


class FakeNano:
    """NanoStats of the reference + report_tables() as csrc/report.cu computes it."""

    def __init__(self, ns):
        self.ns = ns
        self.skipped_reason, self.number_of_reads = ns.skipped_reason, ns.number_of_reads
        self.minimum_time, self.maximum_time = ns.minimum_time, ns.maximum_time

    def report_tables(self, run_start, interval, slots):
        infos = list(self.ns.nano_info_iterator())
        t_bases, t_reads = [0] * slots, [0] * slots
        t_quals = [[0] * 12 for _ in range(slots)]
        active = [set() for _ in range(slots)]
        speeds, parents = [0] * 81, 0
        edges = []
        for k in range(1, 12):  # the largest error / length whose class is >= k, by bisection on the bit pattern
            lo, hi = 1, 0x7ff0000000000000
            while hi - lo > 1:
                mid = (lo + hi) // 2
                x = np.array([mid], dtype=np.uint64).view(np.float64)[0]
                if round(-10 * math.log10(x)) >= 4 * k:
                    lo = mid
                else:
                    hi = mid
            edges.append(float(np.array([lo], dtype=np.uint64).view(np.float64)[0]))
        per_channel = {}
        for i in infos:
            parents += bool(i.parent_id_hash)
            cls = sum((i.cumulative_error_rate / i.length) <= e for e in edges) if i.length else 0
            if i.start_time:
                slot = (i.start_time - run_start) // interval
                t_bases[slot] += i.length
                t_reads[slot] += 1
                t_quals[slot][cls] += 1
                active[slot].add(i.channel_id)
            b, e = per_channel.get(i.channel_id, (0, 0.0))
            per_channel[i.channel_id] = (b + i.length, e + i.cumulative_error_rate)
            if i.duration:
                speeds[int(min(round(i.length / i.duration), 800) // 10)] += 1
        channels = sorted(per_channel)
        return {"time_bases": t_bases, "time_reads": t_reads, "time_active_channels": [len(a) for a in active],
                "time_qualities": t_quals, "translocation_speed": speeds, "reads_with_parent": parents,
                "channels": channels, "channel_bases": [per_channel[c][0] for c in channels],
                "channel_cumulative_error": [per_channel[c][1] for c in channels]}


@pytest.fixture(scope="module")
def ref_env(tmp_path_factory):
    root = tmp_path_factory.mktemp("refpkg")
    pkg = root / "sequali"
    shutil.copytree(PKG_SRC, pkg)
    ref_dir = os.path.join(H.ROOT, "oracle", "_ref", "sequali")
    for so in os.listdir(ref_dir):
        if so.endswith(".so"):
            shutil.copy(os.path.join(ref_dir, so), pkg / so)
    return dict(os.environ, PYTHONPATH=os.pathsep.join([str(root), SHIM]))


def ragged(rng, lengths):
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    out = io.BytesIO()
    for i, ln in enumerate(lengths):
        out.write(b"@r%d\n" % i + letters[rng.integers(0, 5, int(ln))].tobytes() + b"\n+\n" +
                  (rng.integers(0, 60, int(ln)).astype(np.uint8) + 33).tobytes() + b"\n")
    return out.getvalue()


CASES = {
    "nanopore": lambda rng: synth.nanopore_fastq(600, mean_length=1500, max_length=30_000, seed=7),
    "illumina": lambda rng: synth.illumina_fastq(3000, 151, seed=8, n_tiles=4),
    "geometric": lambda rng: ragged(rng, rng.geometric(0.02, size=1500)),
    "with_empty_reads": lambda rng: ragged(rng, np.concatenate([np.zeros(100, dtype=int), rng.integers(0, 40, 300)])),
    "one_read": lambda rng: ragged(rng, [77]),
    "long_tail": lambda rng: ragged(rng, np.concatenate([rng.integers(1, 30, 300), [5000, 9000]])),
    "nothing": lambda rng: b"",
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_report_assembly_and_telescoping_formulas(case, ref_env, tmp_path):
    data = CASES[case](np.random.default_rng(11))
    path = tmp_path / "in.fastq"
    path.write_bytes(data)
    proc = subprocess.run([sys.executable, "-c", REFERENCE_SIDE, str(path), "fastq"], env=ref_env, capture_output=True,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    want = json.loads(proc.stdout)
    qc, ns = REF.QCMetrics(), REF.NanoStats()
    for arr in REF.FastqParser(io.BytesIO(data)):
        qc.add_record_array(arr)
        ns.add_record_array(arr)
    metrics = FakeMetrics(qc)
    ranges = report.data_ranges_for(metrics.max_length)
    assert [list(r) for r in ranges] == want["ranges"]
    got = report.qc_metrics_tables(metrics, ranges)
    assert list(got["aggregated_base_matrix"]) == want["aggregated_base_matrix"]
    assert list(got["aggregated_phred_matrix"]) == want["aggregated_phred_matrix"]
    for name in ("summary", "sequence_length_distribution"):
        w = dict((k[0], k[1]) for k in want[name]["d"])
        for key, value in got[name].items():
            assert enc(value) == w[key], (name, key, value, w[key])
    nano = report.nanostats_report(FakeNano(ns))
    w = dict((k[0], k[1]) for k in want["nanostats"]["d"])
    assert set(w) == set(nano)
    for key, value in nano.items():
        assert enc(value) == w[key], key
