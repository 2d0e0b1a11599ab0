"""-m gpu: the CUDA path (through the C ABI, via the reference-shaped Python API)
against the CPU oracle on the same seeded inputs.  Bit-exact for every table;
doubles are compared by bit pattern."""
import io

import numpy as np
import pytest

from tests import helpers as H
from oracle import oracle as orc
from sequali_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sq():
    import sequali_b200
    return sequali_b200


def _metas_equal(arr, recs):
    m = arr._fetch_metas()
    assert len(m) == len(recs)
    for a, b in (("name_off", "name_off"), ("name_len", "name_len"), ("seq_off", "seq_off"),
                 ("seq_len", "seq_len"), ("qual_off", "qual_off"), ("tags_off", "tags_off"),
                 ("tags_len", "tags_len")):
        assert np.array_equal(m[a].astype(np.uint64), recs[b].astype(np.uint64)), a


@pytest.mark.parametrize("n,length,varlen", [(1, 150, False), (3, 7, True), (1000, 150, False),
                                             (5000, 151, True), (20000, 36, False)])
def test_fastq_parse_matches_oracle(sq, n, length, varlen):
    text = synth.illumina_fastq(n, length=length, seed=n, n_tiles=7, variable_length=varlen,
                                adapter_frac=0.0)
    recs, consumed = orc.parse_fastq(text)
    parser = sq.FastqParser(io.BytesIO(text), 1 << 26)
    arrays = list(parser)
    assert len(arrays) == 1
    assert len(arrays[0]) == len(recs) == n
    _metas_equal(arrays[0], recs)
    assert arrays[0].obj == text
    assert arrays[0][n - 1].name() == bytes(
        text[int(recs[-1]["name_off"]):int(recs[-1]["name_off"]) + int(recs[-1]["name_len"])]
    ).decode()


@pytest.mark.parametrize("bufsize", [1000, 4096, 100_000])
def test_fastq_parse_small_buffers(sq, bufsize):
    text = synth.illumina_fastq(300, length=100, seed=5, n_tiles=3)
    recs, _ = orc.parse_fastq(text)
    total = 0
    names = []
    for arr in sq.FastqParser(io.BytesIO(text), bufsize):
        total += len(arr)
        names.append(arr[0].name())
        assert arr[len(arr) - 1].sequence()
    assert total == len(recs)


def _qc_case(sq, text, bufsize=1 << 26, chunk=None):
    recs, _ = orc.parse_fastq(text)
    oq = orc.QCMetrics()
    buf = np.frombuffer(text, np.uint8)
    oq.add(buf, recs)
    gq = sq.QCMetrics()
    arrays = []
    for arr in sq.FastqParser(io.BytesIO(text), bufsize):
        gq.add_record_array(arr)
        arrays.append(arr)
    H.assert_same(H.dump_qc(gq), H.odump_qc(oq))
    # err_sum written back into the arrays (reference :2126)
    got = np.concatenate([a._fetch_metas()["err_sum"] for a in arrays])
    assert np.array_equal(got.view(np.uint64), recs["err_sum"].view(np.uint64))


def test_qc_illumina_fixed_length(sq):
    _qc_case(sq, synth.illumina_fastq(20000, seed=1, n_tiles=10))


def test_qc_illumina_many_batches(sq):
    _qc_case(sq, synth.illumina_fastq(20000, seed=2, n_tiles=10), bufsize=300_000)


def test_qc_variable_length(sq):
    _qc_case(sq, synth.illumina_fastq(20000, length=151, seed=3, n_tiles=10, variable_length=True))


def test_qc_long_reads(sq):
    _qc_case(sq, synth.nanopore_fastq(300, mean_length=6000, max_length=120_000, seed=4))


def test_qc_lowercase_and_iupac(sq):
    rng = np.random.default_rng(9)
    recs = []
    for i in range(500):
        ln = int(rng.integers(0, 300))
        seq = "".join(rng.choice(list("ACGTacgtNnRYKM-.")) for _ in range(ln))
        qual = "".join(chr(int(x)) for x in rng.integers(33, 127, size=ln))
        recs.append(f"@r{i}\n{seq}\n+\n{qual}\n")
    _qc_case(sq, "".join(recs).encode())
