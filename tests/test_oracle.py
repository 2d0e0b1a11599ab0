"""CPU tests (-m "not gpu") that pin the oracle (oracle/sq_oracle.c):

  * against the golden vectors under tests/golden/ (produced by the unmodified
    reference, see tests/golden/make_golden.py) -- always;
  * against the reference itself (oracle/_ref), when it is present, on fresh
    seeded inputs and on the reference's own fixture files when /root/reference
    exists (it does not on the GPU box).
"""
import gzip
import io
import json
import os
import struct

import numpy as np
import pytest

from tests import helpers as H
from oracle import oracle as orc
from sequali_b200 import synth

REF = H.import_reference()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built")
REF_DATA = "/root/reference/tests/data"
needs_ref_data = pytest.mark.skipif(REF is None or not os.path.isdir(REF_DATA),
                                    reason="reference fixtures not available")


def golden(name):
    with gzip.open(os.path.join(H.GOLDEN, name), "rb") as f:
        data = f.read()
    return json.loads(data) if name.endswith(".json.gz") else data


def normalise(obj):
    """JSON round trip (tuples -> lists, int keys stay out of the picture)."""
    return json.loads(json.dumps(obj))


def test_error_table_matches_reference_table():
    assert H.f64_bits(orc.error_table()) == golden("error_rates.json.gz")
    assert H.f64_bits([10 ** -(q / 10) for q in range(94)]) == golden("error_rates.json.gz")


def test_golden_illumina_single_end():
    text = golden("illumina_se.fastq.gz")
    got = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=211)
    H.assert_same(normalise(got), golden("illumina_se.json.gz"))


def test_golden_illumina_ragged_small_tables():
    text = golden("illumina_ragged.fastq.gz")
    got = H.oracle_single_end(
        text, H.ILLUMINA_ADAPTERS, chunk_records=97,
        dedup_kwargs=dict(max_stored_fingerprints=120, front_sequence_offset=64,
                          back_sequence_offset=0),
        overrep_kwargs=dict(max_unique_fragments=400, sample_every=2))
    want = golden("illumina_ragged.json.gz")
    assert want["dedup"]["modulo_bits"] >= 2  # the fixture does exercise escalation
    assert want["overrep"]["collected_unique_fragments"] == 400  # ... and the cap
    H.assert_same(normalise(got), want)


def test_golden_paired():
    got = H.oracle_paired(golden("paired_R1.fastq.gz"), golden("paired_R2.fastq.gz"),
                          chunk_records=173)
    H.assert_same(normalise(got), golden("paired.json.gz"))


def test_golden_nanopore():
    got = H.oracle_single_end(golden("nanopore.fastq.gz"), H.NANOPORE_ADAPTERS, chunk_records=7)
    H.assert_same(normalise(got), golden("nanopore.json.gz"))


def bam_stream(raw: bytes) -> bytes:
    l_text = struct.unpack("<I", raw[4:8])[0]
    pos = 8 + l_text
    n_ref = struct.unpack("<I", raw[pos:pos + 4])[0]
    pos += 4
    for _ in range(n_ref):
        pos += 4 + struct.unpack("<I", raw[pos:pos + 4])[0] + 4
    return raw[pos:]


def oracle_bam_all(raw: bytes, adapters):
    packed, recs, consumed, _ = orc.decode_bam(bam_stream(raw))
    qc, ad, ptq = orc.QCMetrics(), orc.AdapterCounter(adapters), orc.PerTileQuality()
    ov, ns = orc.OverrepresentedSequences(), orc.NanoStats()
    dd = orc.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0)
    qc.add(packed, recs)
    reason = None
    if ptq.add(packed, recs) == 1:
        reason = H._header_repr(packed, recs[ptq.skipped_record])
    ov.add(packed, recs)
    ns.add(packed, recs)
    ad.add(packed, recs)
    dd.add(packed, recs)
    return dict(qc=H.odump_qc(qc), adapters=H.odump_adapters(ad), ptq=H.odump_ptq(ptq, reason),
                overrep=H.odump_overrep(ov), dedup=H.odump_dedup(dd), nano=H.odump_nano(ns))


def test_golden_bam():
    got = oracle_bam_all(golden("nanopore.bam.gz"), H.NANOPORE_ADAPTERS)
    H.assert_same(normalise(got), golden("nanopore_bam.json.gz"))


# ---------------------------------------------------------------------------
# direct comparisons with the compiled reference
# ---------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("seed", [1, 2])
def test_reference_single_end_random(seed):
    text = synth.illumina_fastq(4000, length=75 + seed, seed=seed, n_tiles=17,
                                variable_length=bool(seed & 1), tile_runs=not (seed & 1))
    kw = dict(dedup_kwargs=dict(max_stored_fingerprints=300, front_sequence_offset=64,
                                back_sequence_offset=0),
              overrep_kwargs=dict(max_unique_fragments=900, sample_every=1 + seed))
    H.assert_same(H.api_single_end(REF, text, H.ILLUMINA_ADAPTERS, buffersize=50_000, **kw),
                  H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=313, **kw))


@needs_ref
def test_reference_paired_random():
    t1, t2 = synth.paired_fastq(3000, seed=9, error_rate=0.02)
    H.assert_same(H.api_paired(REF, t1, t2), H.oracle_paired(t1, t2, chunk_records=500))


@needs_ref
def test_reference_hashes():
    rng = np.random.default_rng(5)
    d = REF.DedupEstimator(max_stored_fingerprints=100)
    o = orc.DedupEstimator(max_stored_fingerprints=100)
    seqs = ["".join(rng.choice(list("ACGT")) for _ in range(int(rng.integers(0, 40))))
            for _ in range(500)]
    buf, recs = orc.pack_records([("n", s, "I" * len(s)) for s in seqs])
    for s in seqs:
        d.add_sequence(s)
    o.add(buf, recs)
    assert list(d.duplication_counts()) == o.duplication_counts().tolist()
    for k in (0, 1, 0xDEADBEEF, 2 ** 63 + 12345):
        import ctypes
        h = orc.lib().orc_wang64(ctypes.c_uint64(k))
        assert orc.lib().orc_wang64_inverse(ctypes.c_uint64(h)) == k


@needs_ref
def test_reference_parser_errors_and_boundaries():
    # every prefix of a small file: same records, same consumed bytes
    text = synth.illumina_fastq(5, length=9, seed=3, n_tiles=2)
    for end in range(len(text) + 1):
        piece = text[:end]
        recs, consumed = orc.parse_fastq(piece)
        want = text[:end].count(b"\n") // 4
        assert len(recs) == want
        assert piece[:consumed].count(b"\n") == 4 * want
    for bad, code in ((b"not a record", orc.E_NO_AT), (b"@n\nSEQ\n-\n", orc.E_NO_PLUS),
                      (b"@n\nAGA\n+\nGG\n", orc.E_LEN)):
        with pytest.raises(orc.FastqFormatError) as e:
            orc.parse_fastq(bad)
        assert e.value.code == code
        with pytest.raises(ValueError):
            next(REF.FastqParser(io.BytesIO(bad)))


@needs_ref
def test_reference_mate_names():
    cases = [("same", "same"), ("same1", "same2"), ("same with comments", "same different"),
             ("same1", "same3"), ("differnt", "diferent"), ("same2", "same5"), ("a/1 x", "a/2\ty")]
    for n1, n2 in cases:
        a = REF.FastqRecordArrayView([REF.FastqRecordView(n1, "A", "A")])
        b = REF.FastqRecordArrayView([REF.FastqRecordView(n2, "A", "A")])
        assert a.is_mate(b) == orc.names_are_mates(n1.encode(), n2.encode()), (n1, n2)


@needs_ref_data
@pytest.mark.parametrize("name", ["dorado_nanopore_100reads.bam", "simple.unaligned.bam",
                                  "missing_quals.bam", "test_skip.bam", "secondary_alignment.bam"])
def test_reference_bam_fixtures(name):
    raw = gzip.open(os.path.join(REF_DATA, name)).read()
    got = oracle_bam_all(raw, H.NANOPORE_ADAPTERS)
    want = H.api_single_end(REF, b"", H.NANOPORE_ADAPTERS, fileobj=io.BytesIO(raw), bam=True)
    H.assert_same(got, want)


@needs_ref_data
@pytest.mark.parametrize("name,adapters", [
    ("100_nanopore_reads.fastq.gz", H.NANOPORE_ADAPTERS),
    ("100_illumina_adapters.fastq", H.ILLUMINA_ADAPTERS),
    ("LTB-A-BC001_S1_L003_R1_001.fastq.gz", H.ILLUMINA_ADAPTERS)])
def test_reference_fastq_fixtures(name, adapters):
    path = os.path.join(REF_DATA, name)
    text = (gzip.open(path) if name.endswith(".gz") else open(path, "rb")).read()
    H.assert_same(H.oracle_single_end(text, adapters, chunk_records=33),
                  H.api_single_end(REF, text, adapters))


@needs_ref_data
def test_reference_paired_fixture():
    t1 = gzip.open(os.path.join(REF_DATA, "LTB-A-BC001_S1_L003_R1_001.fastq.gz")).read()
    t2 = gzip.open(os.path.join(REF_DATA, "LTB-A-BC001_S1_L003_R2_001.fastq.gz")).read()
    H.assert_same(H.oracle_paired(t1, t2, chunk_records=100), H.api_paired(REF, t1, t2))
