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


def _tiny_records(n, seed):
    """Records of 0..3 bases: far below 32 bytes each, the one-pass parser's scratch overflows
    and the two-pass path takes over."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    for i in range(n):
        L = int(rng.integers(0, 4))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
        qual = bytes(rng.integers(33, 74, L, dtype=np.uint8))
        out += b"@" + (b"r%d" % i if i % 3 else b"") + b"\n" + seq + b"\n+\n" + qual + b"\n"
    return bytes(out)


@pytest.mark.parametrize("n", [1, 700, 5000, 60000])
def test_fastq_parse_tiny_records_both_paths(sq, n):
    text = _tiny_records(n, n)
    recs, _ = orc.parse_fastq(text)
    arrays = list(sq.FastqParser(io.BytesIO(text), 1 << 26))
    assert sum(len(a) for a in arrays) == n
    _metas_equal(arrays[0], recs[:len(arrays[0])])


def test_fastq_parse_newline_runs(sq):
    # long runs of empty lines inside otherwise valid records: many newlines per 64-byte window
    rec = b"@x\n\n+\n\n"
    text = rec * 40000
    recs, _ = orc.parse_fastq(text)
    arr = next(iter(sq.FastqParser(io.BytesIO(text), 1 << 26)))
    assert len(arr) == 40000
    _metas_equal(arr, recs)


def test_fastq_parse_read_n_clips(sq):
    text = synth.illumina_fastq(3000, length=150, seed=3, n_tiles=5)
    recs, _ = orc.parse_fastq(text)
    parser = sq.FastqParser(io.BytesIO(text), 1 << 26)
    a = parser.read(1000)
    b = parser.read(1999)
    c = parser.read(5)
    assert (len(a), len(b), len(c)) == (1000, 1999, 1)
    assert a[999].name() == bytes(text[int(recs[999]["name_off"]):][:int(recs[999]["name_len"])]).decode()
    assert c[0].sequence() == bytes(text[int(recs[2999]["seq_off"]):][:150]).decode()


def _parser_outcome(mod, text, bufsize):
    """(number of records, None) or (None, (exception type, message))."""
    try:
        return sum(len(a) for a in mod.FastqParser(io.BytesIO(text), bufsize)), None
    except Exception as e:  # noqa: BLE001
        return None, (type(e), str(e))


@pytest.mark.parametrize("where", [0, 1, 777, 2999])
@pytest.mark.parametrize("kind", ["no_at", "no_plus", "length", "ascii", "partial", "partial_short"])
def test_fastq_parse_errors_like_reference(sq, kind, where):
    """Every parser error of the reference (_qcmodule.c:1055-1143), at the first, an inner and the
    last record of a multi-CTA text; message and exception type equal to the reference's."""
    ref = H.import_reference()
    text = bytearray(synth.illumina_fastq(3000, length=150, seed=11, n_tiles=5))
    recs, _ = orc.parse_fastq(bytes(text))
    r = recs[where]
    start = int(r["name_off"]) - 1
    if kind == "no_at":
        text[start] = ord("A")
    elif kind == "no_plus":
        text[int(r["seq_off"]) + 151] = ord("-")
    elif kind == "length":
        del text[int(r["qual_off"]) + 10]
    elif kind == "ascii":
        text[int(r["seq_off"]) + 3] = 0xC3
    elif kind == "partial":
        del text[len(text) - 40:]
    else:
        del text[int(recs[2999]["name_off"]) + 1:]  # "@x": a tail too short to be looked at
    text = bytes(text)
    got = _parser_outcome(sq, text, 1 << 26)
    if ref is not None:
        want = _parser_outcome(ref, text, 1 << 26)
        assert got == want
    else:
        pat = dict(no_at="does not start with @", no_plus="start with +", length="equal length",
                   ascii="ASCII", partial="ncomplete record", partial_short="ncomplete record")[kind]
        assert got[0] is None and pat in got[1][1]


@pytest.mark.parametrize("bufsize", [1, 2, 7, 100])
def test_fastq_parser_tiny_initial_buffersize(sq, bufsize):
    """The argument is a minimum read step (reference tests/test_fastq_parser.py:141-145): any value >= 1
    parses, and the arrays' `obj` concatenate to the text (:148-174)."""
    text = synth.illumina_fastq(40, length=60, seed=9, n_tiles=3)
    recs, _ = orc.parse_fastq(text)
    arrays = list(sq.FastqParser(io.BytesIO(text), bufsize))
    assert sum(len(a) for a in arrays) == len(recs)
    names = [a[i].name() for a in arrays for i in range(len(a))]
    assert names[0] == bytes(text[1:text.index(b"\n")]).decode() and len(set(names)) == len(recs)
    # obj = leftover + what was read for this array (the partial next record included), as in the reference
    assert all(a.obj.startswith(b"@") for a in arrays)
    ref = H.import_reference()
    if ref is not None:
        ref_arrays = list(ref.FastqParser(io.BytesIO(text), bufsize))
        assert [len(a) for a in arrays] == [len(a) for a in ref_arrays]
        assert [a.obj for a in arrays] == [a.obj for a in ref_arrays]


def test_fastq_parser_empty_input_and_read_exactly_n(sq):
    assert list(sq.FastqParser(io.BytesIO(b""), 1024)) == []
    text = synth.illumina_fastq(25, length=50, seed=10, n_tiles=2)
    parser = sq.FastqParser(io.BytesIO(text), 64)
    sizes = [len(parser.read(n)) for n in (1, 7, 10, 100, 5)]
    assert sizes == [1, 7, 10, 7, 0]
    assert parser.read(1).obj == b""


@pytest.mark.parametrize("bufsize", [4, 40, 1000])
def test_bam_parser_tiny_initial_buffersize(sq, bufsize):
    bam = synth.nanopore_ubam(12, mean_length=300, max_length=2000, seed=12)
    stream = bam[len(synth.bam_header()):]
    packed, recs, consumed, skipped = orc.decode_bam(stream)
    n, got = 0, b""
    for arr in sq.BamParser(io.BytesIO(bam), bufsize):
        n += len(arr)
        got += arr.obj
    assert n == len(recs) and got == packed.tobytes()


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


# ----------------------------------------------------------------------------
# adapters / dedup / overrepresented
# ----------------------------------------------------------------------------
def _run_modules(sq, text, adapters, bufsize, mods, dedup_kwargs=None, overrep_kwargs=None):
    """Selected single-end modules through the GPU API and the oracle."""
    recs, _ = orc.parse_fastq(text)
    buf = np.frombuffer(text, np.uint8)
    out_g, out_o = {}, {}
    g = {}
    if "adapters" in mods:
        g["adapters"] = sq.AdapterCounter(adapters)
    if "dedup" in mods:
        g["dedup"] = sq.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=64,
                                                                back_sequence_offset=0)))
    if "overrep" in mods:
        g["overrep"] = sq.OverrepresentedSequences(**(overrep_kwargs or {}))
    for arr in sq.FastqParser(io.BytesIO(text), bufsize):
        for m in g.values():
            m.add_record_array(arr)
    if "adapters" in mods:
        o = orc.AdapterCounter(adapters)
        o.add(buf, recs)
        H.assert_same(H.dump_adapters(g["adapters"]), H.odump_adapters(o))
    if "dedup" in mods:
        o = orc.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=64,
                                                        back_sequence_offset=0)))
        o.add(buf, recs)
        H.assert_same(H.dump_dedup(g["dedup"]), H.odump_dedup(o))
    if "overrep" in mods:
        o = orc.OverrepresentedSequences(**(overrep_kwargs or {}))
        o.add(buf, recs)
        H.assert_same(H.dump_overrep(g["overrep"]), H.odump_overrep(o))


def test_adapters_illumina(sq):
    text = synth.illumina_fastq(20000, seed=11, n_tiles=5, adapter_frac=0.3)
    _run_modules(sq, text, H.ILLUMINA_ADAPTERS, 1 << 26, {"adapters"})


def test_adapters_long_reads_many_words(sq):
    text = synth.nanopore_fastq(300, mean_length=4000, max_length=60000, seed=12)
    _run_modules(sq, text, H.NANOPORE_ADAPTERS, 1 << 26, {"adapters"})


def test_adapters_odd_patterns(sq):
    # long adapters, N letters (match any non-ACGT), lower case, repeats at chunk borders
    adapters = ["A" * 64, "ACGTN", "GATTACAGATTACAGATTACAGATTACA", "nnnn", "T", "CCGGTTAA" * 8]
    rng = np.random.default_rng(13)
    recs = []
    for i in range(400):
        ln = int(rng.integers(0, 700))
        seq = "".join(rng.choice(list("ACGTacgtNX"), p=[.2, .2, .2, .2, .03, .03, .03, .03, .04, .04])
                      for _ in range(ln))
        if ln > 300 and i % 3 == 0:
            at = int(rng.integers(180, 260))
            ins = adapters[i % len(adapters)]
            seq = (seq[:at] + ins + seq[at:])[:max(ln, at + len(ins))]
        recs.append(f"@r{i}\n{seq}\n+\n{'I' * len(seq)}\n")
    _run_modules(sq, "".join(recs).encode(), adapters, 1 << 26, {"adapters"})


def test_dedup_no_escalation(sq):
    text = synth.illumina_fastq(20000, seed=21, n_tiles=5, dup_frac=0.2)
    _run_modules(sq, text, [], 1 << 26, {"dedup"})


@pytest.mark.parametrize("bufsize", [1 << 26, 200_000])
def test_dedup_escalations(sq, bufsize):
    # 200 slots -> several escalations inside and across record arrays
    text = synth.illumina_fastq(30000, seed=22, n_tiles=5, dup_frac=0.3)
    _run_modules(sq, text, [], bufsize, {"dedup"}, dedup_kwargs=dict(max_stored_fingerprints=200))


@pytest.mark.parametrize("slots", [200, 3000])
def test_dedup_escalations_over_compacted_arrays(sq, slots):
    # record arrays of ~5700 reads: from the second array on only the hashes that pass the current
    # mask reach the table kernels (dedup_consume's compaction), and escalations keep happening inside
    text = synth.illumina_fastq(80000, seed=24, n_tiles=5, dup_frac=0.3)
    _run_modules(sq, text, [], 2_000_000, {"dedup"}, dedup_kwargs=dict(max_stored_fingerprints=slots))


def test_dedup_short_and_odd_fingerprints(sq):
    text = synth.illumina_fastq(5000, length=40, seed=23, n_tiles=5, variable_length=True, dup_frac=0.3)
    for kw in (dict(max_stored_fingerprints=150, front_sequence_length=3, back_sequence_length=5,
                    front_sequence_offset=2, back_sequence_offset=1),
               dict(max_stored_fingerprints=100, front_sequence_length=20, back_sequence_length=1,
                    front_sequence_offset=0, back_sequence_offset=64)):
        _run_modules(sq, text, [], 1 << 26, {"dedup"}, dedup_kwargs=kw)


def test_overrep_default(sq):
    text = synth.illumina_fastq(20000, seed=31, n_tiles=5, dup_frac=0.2)
    _run_modules(sq, text, [], 1 << 26, {"overrep"})


@pytest.mark.parametrize("bufsize", [1 << 26, 150_000])
@pytest.mark.parametrize("kw", [dict(max_unique_fragments=300, sample_every=2),
                                dict(max_unique_fragments=5000, sample_every=1, fragment_length=7),
                                dict(max_unique_fragments=77, sample_every=3, fragment_length=5,
                                     bases_from_start=12, bases_from_end=30)])
def test_overrep_cap_crossing(sq, kw, bufsize):
    text = synth.illumina_fastq(6000, length=90, seed=32, n_tiles=5, variable_length=True)
    _run_modules(sq, text, [], bufsize, {"overrep"}, overrep_kwargs=kw)


# ----------------------------------------------------------------------------
# whole pipelines, shaped like src/sequali/__main__.py:279-306
# ----------------------------------------------------------------------------
@pytest.mark.parametrize("bufsize", [1 << 26, 250_000])
def test_single_end_all_modules_illumina(sq, bufsize):
    text = synth.illumina_fastq(30000, seed=41, n_tiles=40)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=bufsize)
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=997)
    H.assert_same(got, want)


def test_single_end_random_tiles_variable_length(sq):
    text = synth.illumina_fastq(20000, length=120, seed=42, n_tiles=150, tile_runs=False,
                                variable_length=True)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=400_000,
                           dedup_kwargs=dict(max_stored_fingerprints=500),
                           overrep_kwargs=dict(max_unique_fragments=2000, sample_every=3))
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS,
                               dedup_kwargs=dict(max_stored_fingerprints=500),
                               overrep_kwargs=dict(max_unique_fragments=2000, sample_every=3))
    H.assert_same(got, want)


def test_single_end_nanopore(sq):
    text = synth.nanopore_fastq(400, mean_length=5000, max_length=150_000, seed=43)
    got = H.api_single_end(sq, text, H.NANOPORE_ADAPTERS, buffersize=1 << 21)
    want = H.oracle_single_end(text, H.NANOPORE_ADAPTERS)
    H.assert_same(got, want)


def test_pertile_skips_mid_stream(sq):
    good = synth.illumina_fastq(3000, length=50, seed=44, n_tiles=9)
    bad = b"@not_an_illumina_header\nACGT\n+\nIIII\n"
    more = synth.illumina_fastq(500, length=50, seed=45, n_tiles=9)
    text = good + bad + more
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=60_000)
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS)
    H.assert_same(got, want)


@pytest.mark.parametrize("bufsize", [1 << 26, 300_000])
def test_paired_end(sq, bufsize):
    t1, t2 = synth.paired_fastq(20000, seed=46)
    got = H.api_paired(sq, t1, t2, buffersize=bufsize)
    want = H.oracle_paired(t1, t2)
    H.assert_same(got, want)


def test_paired_adapter_table_cap(sq):
    t1, t2 = synth.paired_fastq(8000, seed=47, error_rate=0.03)
    r1, _ = orc.parse_fastq(t1)
    r2, _ = orc.parse_fastq(t2)
    o = orc.InsertSizeMetrics(max_adapters=150)
    o.add_pair(np.frombuffer(t1, np.uint8), r1, np.frombuffer(t2, np.uint8), r2)
    g = sq.InsertSizeMetrics(max_adapters=150)
    p1, p2 = sq.FastqParser(io.BytesIO(t1), 200_000), sq.FastqParser(io.BytesIO(t2), 200_000)
    for a in p1:
        g.add_record_array_pair(a, p2.read(len(a)))
    H.assert_same(H.dump_insert(g), H.odump_insert(o))


def test_bam_nanopore(sq):
    bam = synth.nanopore_ubam(300, mean_length=3000, max_length=80_000, seed=48)
    stream = bam[len(synth.bam_header()):]
    packed, recs, consumed, skipped = orc.decode_bam(stream)
    assert consumed == len(stream)
    oq, ons, oad = orc.QCMetrics(), orc.NanoStats(), orc.AdapterCounter(H.NANOPORE_ADAPTERS)
    oq.add(packed, recs)
    ons.add(packed, recs)
    oad.add(packed, recs)
    gq, gns, gad = sq.QCMetrics(), sq.NanoStats(), sq.AdapterCounter(H.NANOPORE_ADAPTERS)
    got_bytes = b""
    for arr in sq.BamParser(io.BytesIO(bam), 1 << 20):
        gq.add_record_array(arr)
        gns.add_record_array(arr)
        gad.add_record_array(arr)
        got_bytes += arr.obj
        assert arr[0].tags().startswith(b"qsC")
    assert got_bytes == packed.tobytes()
    H.assert_same(H.dump_qc(gq), H.odump_qc(oq))
    H.assert_same(H.dump_nano(gns), H.odump_nano(ons))
    H.assert_same(H.dump_adapters(gad), H.odump_adapters(oad))


# ----------------------------------------------------------------------------
# PerTileQuality: long ordered chains (many reads per tile), power-of-two
# crossings, round-half ties (phred 0..3 against sums in [1, 4)), several arrays
# ----------------------------------------------------------------------------
def _pertile_case(sq, text, bufsize):
    recs, _ = orc.parse_fastq(text)
    o = orc.PerTileQuality()
    o.add(np.frombuffer(text, np.uint8), recs)
    g = sq.PerTileQuality()
    for arr in sq.FastqParser(io.BytesIO(text), bufsize):
        g.add_record_array(arr)
    H.assert_same(H.dump_ptq(g), H.odump_ptq(o))


def _tile_fastq(n, length, n_tiles, qlo, qhi, seed, runs=True, variable=False):
    rng = np.random.default_rng(seed)
    tiles = synth.novaseq_tiles()[:n_tiles]
    t = rng.integers(0, n_tiles, size=n)
    if runs:
        t = np.sort(t)
    qual = (rng.integers(qlo, qhi + 1, size=(n, length)) + 33).astype(np.uint8)
    seq = np.full((n, length), ord("A"), dtype=np.uint8)
    lens = rng.integers(0, length + 1, size=n) if variable else np.full(n, length)
    out = io.BytesIO()
    for i in range(n):
        ln = int(lens[i])
        out.write(b"@SIM:1:FCX:1:%d:%d:%d 1:N:0:ATCACG\n" % (tiles[t[i]], i, i))
        out.write(seq[i, :ln].tobytes() + b"\n+\n" + qual[i, :ln].tobytes() + b"\n")
    return out.getvalue()


@pytest.mark.parametrize("qlo,qhi", [(2, 41), (0, 3), (30, 93), (0, 93)])
@pytest.mark.parametrize("bufsize", [1 << 27, 3_000_000])
def test_pertile_long_chains(sq, qlo, qhi, bufsize):
    _pertile_case(sq, _tile_fastq(120_000, 24, 3, qlo, qhi, seed=qlo * 100 + qhi), bufsize)


def _outlier_fastq(n, length, seed, n_out):
    """Qualities 30..37 everywhere except n_out single bases with phred 0 / 93 / 12 / 60: values the
    range sample of k_fused_reads will mostly miss (slow path of k_fused_columns, tiles replayed)."""
    rng = np.random.default_rng(seed)
    tiles = synth.novaseq_tiles()[:4]
    t = np.sort(rng.integers(0, 4, size=n))
    qual = (rng.integers(30, 38, size=(n, length)) + 33).astype(np.uint8)
    for k in range(n_out):
        qual[int(rng.integers(0, n)), int(rng.integers(0, length))] = 33 + (0, 93, 12, 60)[k % 4]
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(n, length))]
    out = io.BytesIO()
    for i in range(n):
        out.write(b"@SIM:1:FCX:1:%d:%d:%d 1:N:0:ATCACG\n" % (tiles[t[i]], i, i))
        out.write(seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n")
    return out.getvalue()


@pytest.mark.parametrize("bufsize", [1 << 27, 2_000_000])
@pytest.mark.parametrize("n_out", [1, 40])
def test_quality_outliers_beside_sampled_range(sq, n_out, bufsize):
    text = _outlier_fastq(40_000, 75, seed=n_out, n_out=n_out)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=bufsize)
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS)
    H.assert_same(got, want)
    _pertile_case(sq, text, bufsize)  # PerTileQuality alone (no QCMetrics on the array)


def test_pertile_random_tiles_variable_length(sq):
    _pertile_case(sq, _tile_fastq(60_000, 37, 11, 2, 41, seed=7, runs=False, variable=True), 1_000_000)


# ----------------------------------------------------------------------------
# fused short-read pass (csrc/fused.cu): odd content through the one-walk kernel
# ----------------------------------------------------------------------------
def _odd_short_fastq(n, max_len, seed, good_headers=True):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-.*X", np.uint8)
    probs = np.array([.2, .2, .2, .2, .03, .03, .03, .03, .02, .01] + [.00625] * 8)
    probs /= probs.sum()
    ins = [b"AGATCGGAAGAG", b"ACGTN", b"nnnn", b"T", b"GATTACAGATTACAGATTACAGATTACA", b"CTGTCTCTTATA"]
    tiles = synth.novaseq_tiles()[:9]
    out = io.BytesIO()
    for i in range(n):
        ln = int(rng.integers(0, max_len + 1))
        seq = alphabet[rng.choice(len(alphabet), size=ln, p=probs)].tobytes()
        if ln > 40 and i % 3 == 0:
            a = ins[i % len(ins)]
            at = int(rng.integers(0, ln - len(a) + 1)) if ln >= len(a) else 0
            seq = (seq[:at] + a + seq[at + len(a):])[:ln]
        qual = (rng.integers(0, 94, size=ln) + 33).astype(np.uint8).tobytes()
        if good_headers:
            name = b"SIM:1:FCX:1:%d:%d:%d 1:N:0:ATCACG" % (tiles[int(rng.integers(0, 9))], i, i * 7)
        else:
            name = [b"r%d" % i, b"a:b:c:d:%d:x" % i, b"a:b:c:d::x", b"a:b:c:d:12", b"::::7:",
                    b"a:b:c:d:1234567890123456789:x", b"a:b:c:d:12a:x"][i % 7]
        out.write(b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n")
    return out.getvalue()


FUSED_ADAPTERS = ["AGATCGGAAGAG", "ACGTN", "nnnn", "T", "GATTACAGATTACAGATTACAGATTACA", "A" * 32,
                  "CTGTCTCTTATA"]


@pytest.mark.parametrize("max_len", [40, 150, 250, 320])
@pytest.mark.parametrize("bufsize", [1 << 26, 70_000])
def test_fused_odd_content(sq, max_len, bufsize):
    text = _odd_short_fastq(6000, max_len, seed=max_len)
    kw = dict(dedup_kwargs=dict(max_stored_fingerprints=700),
              overrep_kwargs=dict(max_unique_fragments=3000, sample_every=2))
    got = H.api_single_end(sq, text, FUSED_ADAPTERS, buffersize=bufsize, **kw)
    want = H.oracle_single_end(text, FUSED_ADAPTERS, **kw)
    H.assert_same(got, want)


def test_fused_header_variants_switch_pertile_off(sq):
    good = _odd_short_fastq(700, 100, seed=5)
    text = good + _odd_short_fastq(300, 100, seed=6, good_headers=False)
    got = H.api_single_end(sq, text, FUSED_ADAPTERS, buffersize=50_000)
    want = H.oracle_single_end(text, FUSED_ADAPTERS)
    H.assert_same(got, want)


def test_fused_invalid_phred_raises(sq):
    text = b"@SIM:1:FCX:1:1101:1:1 1:N:0:A\nACGTACGTAC\n+\nIIII\x1fIIIII\n" * 3
    qc = sq.QCMetrics()
    for arr in sq.FastqParser(io.BytesIO(text), 1 << 20):
        qc.add_record_array(arr)
    with pytest.raises(ValueError, match="Not a valid phred character"):
        qc.base_count_table()


def test_deferred_adds_keep_call_order(sq):
    # the same collector fed with two arrays before any getter, and NanoStats fed before
    # QCMetrics (it must then see error sums of 0.0, like the reference: _qcmodule.c:5314)
    text = synth.nanopore_fastq(40, mean_length=200, max_length=300, seed=3)
    recs, _ = orc.parse_fastq(text)
    buf = np.frombuffer(text, np.uint8)
    ons, oq = orc.NanoStats(), orc.QCMetrics()
    gns, gq = sq.NanoStats(), sq.QCMetrics()
    arrays = list(sq.FastqParser(io.BytesIO(text), 3000))
    assert len(arrays) > 2
    start = 0
    for arr in arrays:
        part = recs[start:start + len(arr)]
        start += len(arr)
        ons.add(buf, part)
        oq.add(buf, part)
        gns.add_record_array(arr)
        gq.add_record_array(arr)
    H.assert_same(H.dump_nano(gns), H.odump_nano(ons))
    H.assert_same(H.dump_qc(gq), H.odump_qc(oq))


# ----------------------------------------------------------------------------
# the bench input itself (device-generated NovaSeq recipe C2), at a size that
# crosses the 5 M unique-fragment cap of OverrepresentedSequences and makes
# DedupEstimator escalate, through the HBM-resident path bench.py times
# ----------------------------------------------------------------------------
def test_bench_input_prefix_all_modules(sq):
    from sequali_b200.device import DeviceFastq
    n = 6_500_000
    data = DeviceFastq.synth_illumina(n, 150, seed=2, chunk_reads=1 << 21, total_reads=100_000_000)
    mods = dict(qc=sq.QCMetrics(), ptq=sq.PerTileQuality(), ov=sq.OverrepresentedSequences(),
                ns=sq.NanoStats(), ad=sq.AdapterCounter(H.ILLUMINA_ADAPTERS),
                dd=sq.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0))
    for arr in data.record_arrays():
        for key in ("qc", "ptq", "ov", "ns", "ad", "dd"):
            mods[key].add_record_array(arr)
    got = dict(qc=H.dump_qc(mods["qc"]), adapters=H.dump_adapters(mods["ad"]), ptq=H.dump_ptq(mods["ptq"]),
               overrep=H.dump_overrep(mods["ov"]), dedup=H.dump_dedup(mods["dd"]), nano=H.dump_nano(mods["ns"]))
    host, reads = data.to_host()
    assert reads == n
    want = H.oracle_single_end(host.tobytes(), H.ILLUMINA_ADAPTERS, chunk_records=1 << 20)
    assert want["overrep"]["collected_unique_fragments"] == 5_000_000   # the cap was hit
    assert want["dedup"]["modulo_bits"] >= 1                               # at least one escalation
    H.assert_same(got, want)


# ----------------------------------------------------------------------------
# the host reader bench.py's e2e leg uses (sq_fastq_stream_*): windows of pinned
# host text copied ahead of the parser, leftovers carried on the device
# ----------------------------------------------------------------------------
def _host_stream_case(sq, text, window):
    import ctypes
    from sequali_b200.device import HostFastq
    from sequali_b200._lib import Context
    ctx = Context.get()
    hq = HostFastq(ctx, len(text))
    ctypes.memmove(hq.ptr, text, len(text))
    mods = dict(qc=sq.QCMetrics(), ptq=sq.PerTileQuality(), ov=sq.OverrepresentedSequences(),
                ns=sq.NanoStats(), ad=sq.AdapterCounter(H.ILLUMINA_ADAPTERS),
                dd=sq.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0))
    n, arrays = 0, 0
    for arr in hq.record_arrays(window):
        n += len(arr)
        arrays += 1
        for key in ("qc", "ptq", "ov", "ns", "ad", "dd"):
            mods[key].add_record_array(arr)
    got = dict(qc=H.dump_qc(mods["qc"]), adapters=H.dump_adapters(mods["ad"]), ptq=H.dump_ptq(mods["ptq"]),
               overrep=H.dump_overrep(mods["ov"]), dedup=H.dump_dedup(mods["dd"]), nano=H.dump_nano(mods["ns"]))
    hq.free()
    return got, n, arrays


@pytest.mark.parametrize("window", [4096, 65_536, 1 << 26])
def test_host_stream_windows(sq, window):
    text = synth.illumina_fastq(3000, length=150, seed=31, n_tiles=9)
    got, n, arrays = _host_stream_case(sq, text, window)
    assert n == 3000 and arrays == min(3000, -(-len(text) // window))
    H.assert_same(got, H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=700))


def test_host_stream_record_longer_than_window(sq):
    # nanopore-sized reads: several windows hold no complete record and are joined with the next
    text = synth.nanopore_fastq(40, seed=12, mean_length=9000)
    recs, _ = orc.parse_fastq(text)
    assert max(int(r["seq_len"]) for r in recs) > 8192
    got, n, _ = _host_stream_case(sq, text, 4096)
    assert n == len(recs)
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS)
    H.assert_same(got, want)


def test_host_stream_partial_tail_raises(sq):
    text = synth.illumina_fastq(50, length=100, seed=3, n_tiles=2)
    with pytest.raises(EOFError):
        _host_stream_case(sq, text[:-7], 4096)
