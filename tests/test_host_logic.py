"""CPU tests of the host-side mirror of the reference interface: argument
checking and error text that the reference's own tests pin and
that do not need a device."""
import pytest

import sequali_b200 as sq
from sequali_b200 import _qc


def test_constants_match_reference_module():
    assert (sq.A, sq.C, sq.G, sq.T, sq.N) == (0, 1, 2, 3, 4)
    assert sq.NUMBER_OF_NUCS == 5 and sq.NUMBER_OF_PHREDS == 12 and sq.TABLE_SIZE == 60
    assert sq.PHRED_MAX == 93 and sq.MAX_SEQUENCE_SIZE == 64
    assert sq.DEFAULT_MAX_UNIQUE_FRAGMENTS == 5_000_000
    assert sq.DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS == 1_000_000
    assert sq.DEFAULT_FRAGMENT_LENGTH == 21 and sq.DEFAULT_UNIQUE_SAMPLE_EVERY == 8
    assert sq.INSERT_SIZE_MAX_ADAPTER_STORE_SIZE == 31 and sq.DEFAULT_END_ANCHOR_LENGTH == 100


def test_record_view_accessors_and_validation():
    v = sq.FastqRecordView("name x", "ACGT", "IIII", b"RGZA\x00")
    assert (v.name(), v.sequence(), v.qualities(), v.tags()) == ("name x", "ACGT", "IIII", b"RGZA\x00")
    assert v.obj == b"name xACGTIIIIRGZA\x00"
    with pytest.raises(ValueError, match="different lengths"):
        sq.FastqRecordView("n", "ACGT", "III")
    with pytest.raises(ValueError, match="ASCII"):
        sq.FastqRecordView("nä", "A", "I")
    with pytest.raises(ValueError, match="Not a valid phred character"):
        sq.FastqRecordView("n", "A", " ")
    with pytest.raises(TypeError):
        sq.FastqRecordView(b"n", "A", "I")


def test_record_array_from_views():
    views = [sq.FastqRecordView(f"r{i}", "ACGT" * i, "I" * (4 * i)) for i in range(4)]
    arr = sq.FastqRecordArrayView(views)
    assert len(arr) == 4 and arr[2].sequence() == "ACGTACGT" and arr[-1].name() == "r3"
    assert arr.obj == b"".join(v.obj for v in views)
    with pytest.raises(IndexError):
        arr[4]
    with pytest.raises(TypeError, match="FastqRecordView"):
        sq.FastqRecordArrayView([b"x"])
    with pytest.raises(TypeError, match="FastqRecordArrayView"):
        arr.is_mate("nope")


@pytest.mark.parametrize("make,exc,pattern", [
    (lambda: sq.AdapterCounter([]), ValueError, "t least one"),
    (lambda: sq.AdapterCounter(1), TypeError, "not iterable"),
    (lambda: sq.AdapterCounter(["GATTACA", b"GATTACA"]), TypeError, "b'GATTACA'"),
    (lambda: sq.AdapterCounter(["GATTACA", "Gättaca"]), ValueError, "ASCII"),
    (lambda: sq.AdapterCounter(["A" * 31, "A" * 65]), ValueError, "65"),
    (lambda: sq.QCMetrics(end_anchor_length=-1), ValueError, "end_anchor_length"),
    (lambda: sq.OverrepresentedSequences(fragment_length=4), ValueError, "uneven"),
    (lambda: sq.OverrepresentedSequences(sample_every=0), ValueError, "sample_every"),
    (lambda: sq.OverrepresentedSequences(max_unique_fragments=0), ValueError, "at least 1"),
    (lambda: sq.DedupEstimator(max_stored_fingerprints=7), ValueError, "max_stored_fingerprints"),
    (lambda: sq.DedupEstimator(front_sequence_length=-1), ValueError, "front_sequence_length"),
    (lambda: sq.DedupEstimator(back_sequence_offset=-1), ValueError, "back_sequence_offset"),
    (lambda: sq.InsertSizeMetrics(max_adapters=0), ValueError, "max_adapters"),
    (lambda: sq.FastqParser(None, initial_buffersize=0), ValueError, "at least 1"),
])
def test_constructor_argument_errors(make, exc, pattern):
    with pytest.raises(exc, match=pattern):
        make()


def test_bam_parser_header_errors():
    import io
    with pytest.raises(ValueError, match="at least 4"):
        sq.BamParser(io.BytesIO(), initial_buffersize=3)
    with pytest.raises(ValueError, match="not a BAM"):
        sq.BamParser(io.BytesIO(b"@my header"))
    with pytest.raises(EOFError, match="runcated BAM"):
        sq.BamParser(io.BytesIO(b"BAM\x01\x10\x00\x00\x00abc"))
    with pytest.raises(TypeError, match="binary IO"):
        sq.BamParser(io.StringIO("BAM\x01...."))


def test_kmer_helper_round_trip():
    assert _qc._kmer_to_sequence(0b00011011, 4) == "ACGT"
