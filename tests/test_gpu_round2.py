"""-m gpu: cases added in round 2 -- limits lifted, documented behavioural differences pinned."""
import io
import warnings

import numpy as np
import pytest

from sequali_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

REF = H.import_reference()


@pytest.fixture(scope="module")
def sq():
    import sequali_b200
    return sequali_b200


@pytest.mark.skipif(REF is None, reason="oracle/_ref is not built")
@pytest.mark.parametrize("kwargs", [dict(bases_from_start=-1, bases_from_end=-1),
                                    dict(bases_from_start=5000, bases_from_end=-1, sample_every=1),
                                    dict(bases_from_start=-1, bases_from_end=0, fragment_length=31, sample_every=2)])
def test_overrepresented_whole_read_fragments_of_long_reads(sq, kwargs):
    """_qcmodule.c:3499-3504: negative bases_from_* = the whole read; thousands of fragments per read
    (the per-read staging set then lives in the fragment buffer itself, not in local memory)."""
    text = synth.nanopore_fastq(60, mean_length=9000, max_length=60_000, seed=77)

    def run(mod):
        ov = mod.OverrepresentedSequences(**kwargs)
        for arr in mod.FastqParser(io.BytesIO(text), 1 << 20):
            ov.add_record_array(arr)
        return H.dump_overrep(ov)
    H.assert_same(run(sq), run(REF))


def _fastq(names, seqs):
    return b"".join(b"@" + n + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for n, s in zip(names, seqs))


@pytest.mark.skipif(REF is None, reason="oracle/_ref is not built")
def test_dedup_pair_stale_bytes_across_batches(sq):
    """_qcmodule.c:4503-4516: a pair shorter than the fingerprint hashes what the PREVIOUS pair left in
    the estimator's scratch buffer -- also when that pair belonged to an earlier record array."""
    rng = np.random.default_rng(5)
    n = 400
    seqs1, seqs2 = [], []
    for i in range(n):
        # arrays 0 and 2 (100 pairs each) hold no short pair; arrays 1 and 3 start with one / hold several
        short = (100 <= i < 200 and i % 17 in (0, 1)) or i == 300 or i == 399
        l1, l2 = (int(rng.integers(0, 8)), int(rng.integers(0, 8))) if short else (int(rng.integers(8, 60)),) * 2
        seqs1.append(bytes(rng.choice(list(b"ACGT"), l1).astype(np.uint8)))
        seqs2.append(bytes(rng.choice(list(b"ACGT"), l2).astype(np.uint8)))
    names = [b"r%d" % i for i in range(n)]
    t1, t2 = _fastq(names, seqs1), _fastq(names, seqs2)

    def run(mod):
        dd = mod.DedupEstimator(front_sequence_offset=0, back_sequence_offset=0)
        p1, p2 = mod.FastqParser(io.BytesIO(t1), 1 << 20), mod.FastqParser(io.BytesIO(t2), 1 << 20)
        while True:
            a = p1.read(100)
            if len(a) == 0:
                break
            dd.add_record_array_pair(a, p2.read(len(a)))
        return H.dump_dedup(dd)
    H.assert_same(run(sq), run(REF))


# ---- the CPython extension (sequali_b200/ext/qc_ext.cpp) on whole runs --------------------------------
@pytest.fixture(scope="module")
def ext():
    import sequali_b200.ext
    return sequali_b200.ext


def test_extension_single_end_illumina(ext):
    text = synth.illumina_fastq(30_000, length=150, seed=21, n_tiles=24)
    for bufsize in (100_000, 4 << 20):
        H.assert_same(H.api_single_end(ext, text, H.ILLUMINA_ADAPTERS, buffersize=bufsize),
                      H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=777))


def test_extension_paired(ext):
    t1, t2 = synth.paired_fastq(6000, seed=22)
    H.assert_same(H.api_paired(ext, t1, t2, buffersize=300_000), H.oracle_paired(t1, t2))


def test_extension_nanopore_fastq_and_bam(ext, sq):
    text = synth.nanopore_fastq(200, mean_length=4000, max_length=60_000, seed=23)
    H.assert_same(H.api_single_end(ext, text, H.NANOPORE_ADAPTERS, buffersize=1 << 20),
                  H.oracle_single_end(text, H.NANOPORE_ADAPTERS, chunk_records=50))
    bam = synth.nanopore_ubam(200, mean_length=3000, max_length=50_000, seed=24)
    H.assert_same(H.api_single_end(ext, b"", H.NANOPORE_ADAPTERS, buffersize=1 << 20, fileobj=io.BytesIO(bam), bam=True),
                  H.api_single_end(sq, b"", H.NANOPORE_ADAPTERS, buffersize=1 << 20, fileobj=io.BytesIO(bam), bam=True))


def test_extension_parser_errors_and_record_access(ext, sq):
    text = synth.illumina_fastq(50, length=40, seed=25, n_tiles=3)
    for bad in (text.replace(b"\n+\n", b"\n-\n", 1), b"x" + text, text[:-7], text.replace(b"A", b"\xc3", 1),
                text[:200] + text[203:]):
        outs = []
        for mod in (ext, sq):
            try:
                outs.append([len(a) for a in mod.FastqParser(io.BytesIO(bad), 997)])
            except Exception as e:  # noqa: BLE001
                outs.append((type(e).__name__, str(e)))
        assert outs[0] == outs[1], outs
    arrs = [list(m.FastqParser(io.BytesIO(text), 1500)) for m in (ext, sq)]
    assert [len(a) for a in arrs[0]] == [len(a) for a in arrs[1]]
    for a, b in zip(*arrs):
        assert a.obj == b.obj
        for i in (0, len(a) - 1):
            assert (a[i].name(), a[i].sequence(), a[i].qualities(), a[i].tags()) == \
                   (b[i].name(), b[i].sequence(), b[i].qualities(), b[i].tags())


# ---- BGZF members inflated on the device (sq_fastq_stream_create_bgzf, csrc/inflate.cu) ----------------
@pytest.mark.parametrize("level,block_text,window", [(6, 65280, 1 << 20), (1, 20000, 300_000), (0, 60000, 65536),
                                                     (9, 65280, 64 << 20)])
def test_bgzf_fastq_stream_equals_plain_text(sq, level, block_text, window):
    from sequali_b200.device import HostFastq
    text = synth.illumina_fastq(12_000, length=150, seed=31, n_tiles=9)
    comp = synth.bgzf_compress(text, level=level, block_text=block_text)
    host = HostFastq.from_bytes(comp)
    qc, ad, ptq = sq.QCMetrics(), sq.AdapterCounter(H.ILLUMINA_ADAPTERS), sq.PerTileQuality()
    ov, ns = sq.OverrepresentedSequences(), sq.NanoStats()
    dd = sq.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0)
    n, got_text = 0, []
    for arr in host.record_arrays_bgzf(window):
        n += len(arr)
        got_text.append(arr.obj[:])  # the inflated bytes of this array (leftover of the previous one in front)
        for m in (qc, ptq, ov, ns, ad, dd):
            m.add_record_array(arr)
    got = dict(qc=H.dump_qc(qc), adapters=H.dump_adapters(ad), ptq=H.dump_ptq(ptq), overrep=H.dump_overrep(ov),
               dedup=H.dump_dedup(dd), nano=H.dump_nano(ns))
    assert n == 12_000
    H.assert_same(got, H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=2500))
    host.free()


def test_bgzf_corrupt_member_is_an_error(sq):
    from sequali_b200.device import HostFastq
    text = synth.illumina_fastq(3000, length=150, seed=32, n_tiles=3)
    comp = bytearray(synth.bgzf_compress(text, level=6, block_text=30000))
    comp[len(comp) // 2] ^= 0x10  # inside the payload of a member in the middle
    host = HostFastq.from_bytes(bytes(comp))
    with pytest.raises((ValueError, EOFError)):
        for _ in host.record_arrays_bgzf(1 << 20):
            pass
    host.free()
    with pytest.raises(ValueError, match="BGZF"):  # plain gzip has no member index
        import gzip
        h2 = HostFastq.from_bytes(gzip.compress(text))
        list(h2.record_arrays_bgzf(1 << 20))


def test_extension_read_ahead_keeps_the_record_stream(ext, sq):
    """The extension's parsers read one array ahead on a helper thread (read steps >= 1 MiB); mixing
    __next__ and read(n) must still hand out every record exactly once and in order, errors at their place."""
    text = synth.illumina_fastq(30_000, length=100, seed=41, n_tiles=5)  # ~7 MB: several 1 MiB steps

    def walk(mod):
        # (how many records an array holds is the parser's business; the record stream is not)
        p, out = mod.FastqParser(io.BytesIO(text), 1 << 20), []

        def take(arr):
            out.extend(arr[i].name() for i in range(len(arr)))
            return len(arr)
        take(next(p))
        assert take(p.read(7)) == 7         # un-reads what was read ahead
        take(next(p))
        take(p.read(50_000))                # everything that is left
        assert len(p.read(1)) == 0
        return out
    got, want = walk(ext), walk(sq)
    assert len(got) == 30_000 and got == want
    assert sum(len(a) for a in ext.FastqParser(io.BytesIO(text), 1 << 20)) == 30_000
    bad = text[:3_000_000] + b"oops\n" + text[3_000_000:]
    got = []
    for mod in (ext, sq):
        n = 0
        try:
            for arr in mod.FastqParser(io.BytesIO(bad), 1 << 20):
                n += len(arr)
        except ValueError as e:
            got.append((n, str(e)))
    assert len(got) == 2 and got[0] == got[1]
    bam = synth.nanopore_ubam(400, mean_length=8000, max_length=100_000, seed=42)
    assert [len(a) for a in ext.BamParser(io.BytesIO(bam), 1 << 20)] == [len(a) for a in sq.BamParser(io.BytesIO(bam), 1 << 20)]


def _impl(name):
    if name == "extension":
        import sequali_b200.ext
        return sequali_b200.ext
    import sequali_b200
    return sequali_b200


# ---- NanoStats: malformed aux data raises what the reference raises (_qcmodule.c:5078-5259) ----
@pytest.mark.parametrize("impl", ["ctypes", "extension"])
@pytest.mark.parametrize("tags", [b"xxX\1", b"xxBZ\1\0\0\0a\0", b"sti\1\0\0\0", b"duZabc\0", b"pii\1\0\0\0",
                                  b"chf\0\0\0\0", b"chi\1\0", b"xxZabc", b"xxBi\xff\0\0\0"])
def test_nanostats_tag_errors_are_the_references(impl, tags):
    import struct  # noqa: F401
    import warnings
    from sequali_b200 import synth
    ref = H.import_reference()
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    sq = _impl(impl)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)
    good = b"chS\x05\x00" + b"stZ2024-01-01T00:00:00Z\0"
    raw = synth.bam_header() + b"".join(
        synth.bam_record(b"r%d" % i, seq, np.full(4, 50, np.uint8), tags if i == 3 else good) for i in range(6))
    outcomes = []
    for mod in (ref, sq):
        ns = mod.NanoStats()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                for arr in mod.BamParser(io.BytesIO(raw)):
                    ns.add_record_array(arr)
                outcomes.append(("ok", ns.number_of_reads))
            except BaseException as e:  # noqa: BLE001
                # (the text of a SystemError is CPython's, and depends on how the method was called)
                outcomes.append((type(e).__name__, "" if isinstance(e, SystemError) else str(e)))
    assert outcomes[0] == outcomes[1] and outcomes[0][0] != "ok"


@pytest.mark.parametrize("impl", ["ctypes", "extension"])
def test_nanostats_pi_warning_names_the_length(impl):
    import warnings
    from sequali_b200 import synth
    sq = _impl(impl)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)
    raw = synth.bam_header() + b"".join(
        synth.bam_record(b"r%d" % i, seq, np.full(4, 50, np.uint8), b"chS\x05\x00" + (b"piZabcdef\0" if i == 2 else b""))
        for i in range(4))
    ns = sq.NanoStats()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for arr in sq.BamParser(io.BytesIO(raw)):
            ns.add_record_array(arr)
        assert ns.number_of_reads == 4
    assert [str(x.message) for x in w] == ["pi tag should have a valid uuid4 format with 36 characters. Counted 6. Skipping tag."]


@pytest.mark.parametrize("impl", ["ctypes", "extension"])
def test_nanostats_skipped_reason_is_learned_without_waiting(impl):
    """Record arrays keep arriving after the header that switches NanoStats off: the reason names that header and
    the counters stop there, whichever add notices it."""
    from sequali_b200 import synth
    ref = H.import_reference()
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    sq = _impl(impl)
    nano = synth.nanopore_fastq(300, mean_length=400, max_length=3000, seed=81)
    lines = nano.split(b"\n")
    lines[4 * 170] = b"@not a nanopore header"
    text = b"\n".join(lines)
    got = []
    for mod in (ref, sq):
        ns = mod.NanoStats()
        for arr in mod.FastqParser(io.BytesIO(text), 20_000):
            ns.add_record_array(arr)
        got.append((ns.skipped_reason, ns.number_of_reads, ns.minimum_time, ns.maximum_time))
    assert got[0] == got[1] and "not a nanopore header" in got[0][0]


# ---- regular files: the extension reads them with pread() from several threads (qc_ext.cpp parser_read) ----
def _names(arrays):
    return [arr[i].name() for arr in arrays for i in range(len(arr))]


def test_extension_reads_regular_files_directly(tmp_path):
    """open(path, 'rb') (BufferedReader), buffering=0 (FileIO) and gzip objects (never direct: their fileno() is
    the compressed file's) give the record stream of an in-memory object; the file object ends up where the
    parser stopped reading."""
    import gzip
    import sequali_b200.ext as ext
    text = synth.illumina_fastq(120_000, 151, seed=91, n_tiles=6)       # ~41 MB: several direct reads of 8 MiB
    want = _names(ext.FastqParser(io.BytesIO(text), 8 << 20))
    assert len(want) == 120_000
    path = tmp_path / "reads.fastq"
    path.write_bytes(text)
    for buffering in (-1, 0, 1 << 16):
        with open(path, "rb", buffering=buffering) as f:
            assert _names(ext.FastqParser(f, 8 << 20)) == want
            assert f.tell() == len(text) and f.read(1) == b""
    gz = tmp_path / "reads.fastq.gz"
    with gzip.open(gz, "wb", compresslevel=1) as f:
        f.write(text)
    with gzip.open(gz, "rb") as f:
        assert _names(ext.FastqParser(f, 8 << 20)) == want
    # read(n) and iteration mixed on a regular file; a partial last record is still an error
    with open(path, "rb") as f:
        p = ext.FastqParser(f, 8 << 20)
        first = p.read(1000)
        rest = _names(p)
        assert _names([first]) + rest == want
    (tmp_path / "cut.fastq").write_bytes(text[:-7])
    with open(tmp_path / "cut.fastq", "rb") as f, pytest.raises(EOFError):
        _names(ext.FastqParser(f, 8 << 20))


def test_extension_file_reader_with_records_longer_than_a_block(tmp_path):
    """Reads of up to 9 Mb with 4 MiB blocks: blocks without a whole record, leftovers longer than the room in front
    of a block; and a parser that is dropped half way leaves the file where its last record array ended."""
    import sequali_b200.ext as ext
    rng = np.random.default_rng(93)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = io.BytesIO()
    for i, ln in enumerate([300, 9_000_000, 50, 2_000_000, 5_000_000, 120, 4_194_304, 7, 1_100_000, 60] * 2):
        out.write(b"@read%d ch=%d start_time=2024-01-01T00:00:%02dZ\n" % (i, i + 1, i))
        out.write(letters[rng.integers(0, 4, ln)].tobytes() + b"\n+\n" + (rng.integers(0, 40, ln).astype(np.uint8) + 33).tobytes() + b"\n")
    text = out.getvalue()
    want = [(a[i].name(), len(a[i].sequence()), a[i].sequence()[:30], a[i].qualities()[-30:])
            for a in ext.FastqParser(io.BytesIO(text), 4 << 20) for i in range(len(a))]
    assert len(want) == 20
    path = tmp_path / "long.fastq"
    path.write_bytes(text)
    with open(path, "rb") as f:
        got = [(a[i].name(), len(a[i].sequence()), a[i].sequence()[:30], a[i].qualities()[-30:])
               for a in ext.FastqParser(f, 4 << 20) for i in range(len(a))]
        assert got == want and f.tell() == len(text)
    with open(path, "rb") as f:
        parser = ext.FastqParser(f, 4 << 20)
        first = next(parser)
        used = len(first.obj)
        del parser, first
        # (the array read ahead on the helper thread holds the parser until it is done; then the reader stops and
        # moves the file object behind the last block it handed out)
        import time
        for _ in range(200):
            if f.tell():
                break
            time.sleep(0.01)
        assert 0 < f.tell() <= len(text) and f.tell() % (4 << 20) == 0 and f.tell() >= used


def test_extension_reads_regular_bam_files_directly(tmp_path):
    """BamParser has consumed the header through read(): the descriptor's own position is somewhere else than the
    object's, the direct reads start at tell()."""
    import sequali_b200.ext as ext
    raw = synth.nanopore_ubam(3000, mean_length=4000, max_length=60_000, seed=92)   # ~20 MB
    assert len(raw) > (12 << 20)

    def dump(fileobj, size):
        parser = ext.BamParser(fileobj, size)
        return parser.header, [(a[i].name(), a[i].sequence()[:50], len(a[i].qualities()), a[i].tags()[:20])
                               for a in parser for i in range(len(a))]

    want = dump(io.BytesIO(raw), 4 << 20)
    assert len(want[1]) == 3000
    path = tmp_path / "reads.bam"
    path.write_bytes(raw)
    for buffering in (-1, 0):
        with open(path, "rb", buffering=buffering) as f:
            assert dump(f, 4 << 20) == want
            assert f.tell() == len(raw)
    # a file that ends inside a record; a file whose records are all secondary alignments
    (tmp_path / "cut.bam").write_bytes(raw[:-11])
    with open(tmp_path / "cut.bam", "rb") as f, pytest.raises(EOFError, match="ncomplete record"):
        dump(f, 4 << 20)
    with pytest.raises(EOFError, match="ncomplete record"):
        dump(io.BytesIO(raw[:-11]), 4 << 20)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, 3000)]
    skipped = synth.bam_header() + b"".join(
        synth.bam_record(b"s%d" % i, seq, np.full(3000, 60, np.uint8), b"", flag=0x100) for i in range(2500))
    (tmp_path / "skipped.bam").write_bytes(skipped)
    with open(tmp_path / "skipped.bam", "rb") as f:
        assert dump(f, 4 << 20)[1] == dump(io.BytesIO(skipped), 4 << 20)[1] == []


def test_large_scratch_blocks_are_reused_after_the_first_passes():
    """Steady state: passes over the same input make no driver allocation of 8 MiB or more any more (sq_dalloc's
    block cache), and the results do not depend on which pass it is."""
    import ctypes as C
    import sequali_b200 as sq
    from sequali_b200 import _lib
    from sequali_b200.device import DeviceFastq
    ctx = _lib.Context.get()
    data = DeviceFastq.synth_illumina(3_000_000, 151, seed=5, chunk_reads=1 << 20)

    def stats():
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(ctx.lib.sq_ctx_block_cache_stats(ctx.h, C.byref(a), C.byref(b), C.byref(c)), "stats")
        return a.value, b.value, c.value

    def one_pass():
        qc, pt, ov, dd = sq.QCMetrics(), sq.PerTileQuality(), sq.OverrepresentedSequences(), sq.DedupEstimator()
        ad = sq.AdapterCounter(["AGATCGGAAGAGC", "CTGTCTCTTATA"])
        for arr in data.record_arrays():
            for m in (qc, pt, ov, ad, dd):
                m.add_record_array(arr)
        return (bytes(qc.base_count_table()), bytes(qc.phred_count_table()), pt.get_tile_counts(),
                bytes(dd.duplication_counts()), ov.sequence_counts(), ad.get_counts())

    first = one_pass()
    one_pass()
    one_pass()
    ctx.sync()
    before = stats()
    again = one_pass()
    ctx.sync()
    after = stats()
    assert again == first
    assert after[0] == before[0], (before, after)      # nothing new from the driver
    assert after[1] > before[1] and after[2] > 0       # the pass lived off the cache
    data.free()


@pytest.mark.parametrize("which", ["both files", "read() side a file", "iterated side a file"])
def test_extension_paired_reads_from_regular_files(tmp_path, which):
    """The paired loop of the CLI: one parser iterated, the other asked for exactly as many records with read(n)."""
    import sequali_b200.ext as ext
    t1, t2 = synth.paired_fastq(60_000, seed=94)                      # ~21 MB each
    p1, p2 = tmp_path / "r1.fastq", tmp_path / "r2.fastq"
    p1.write_bytes(t1)
    p2.write_bytes(t2)

    def run(f1, f2):
        names = []
        rd1, rd2 = ext.FastqParser(f1, 4 << 20), ext.FastqParser(f2, 4 << 20)
        for a in rd1:
            b = rd2.read(len(a))
            assert len(b) == len(a) and a.is_mate(b)
            names.append((a[0].name(), b[len(b) - 1].name(), len(a)))
        assert len(rd2.read(1)) == 0
        return names

    want = run(io.BytesIO(t1), io.BytesIO(t2))
    assert sum(n for _, _, n in want) == 60_000
    with open(p1, "rb") as f1, open(p2, "rb") as f2:
        got = run(f1 if which != "read() side a file" else io.BytesIO(t1),
                  f2 if which != "iterated side a file" else io.BytesIO(t2))
    assert [(a, b) for a, b, _ in got][0] == [(a, b) for a, b, _ in want][0]
    assert sum(n for _, _, n in got) == 60_000
