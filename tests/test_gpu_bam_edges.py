"""-m gpu: BamParser edge cases on the CUDA path (reference tests/test_bam_parser.py:95-147 and
_qcmodule.c:1623-1694): missing qualities (0xff block -> '!'), secondary / supplementary records
skipped, every truncation point of header and records, tiny read steps -- the unmodified reference
extension (oracle/_ref) and sequali_b200 side by side on the same bytes."""
import gzip
import io
import os
import struct

import numpy as np
import pytest

from sequali_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

REF = H.import_reference()
if REF is None:  # pragma: no cover
    pytest.skip("oracle/_ref is not built", allow_module_level=True)

REF_DATA = os.path.join(H.ROOT, "oracle", "_ref", "tests", "tests", "data")


@pytest.fixture(scope="module", params=["ctypes", "extension"])
def sq(request):
    """The B200 build behind its two host layers: the ctypes mirror and the CPython extension."""
    if request.param == "extension":
        import sequali_b200.ext
        return sequali_b200.ext
    import sequali_b200
    return sequali_b200


def make_bam(rng, n, flags=(4,), missing_every=0, lengths=(0, 1, 2, 7, 8, 33, 150, 151, 1000)):
    out = io.BytesIO()
    out.write(synth.bam_header())
    letters = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
    for i in range(n):
        ln = int(lengths[i % len(lengths)])
        seq = letters[rng.integers(0, 16, size=ln)]
        qual = (rng.integers(0, 94, size=ln).astype(np.uint8) + 33)
        tags = b"chS" + struct.pack("<H", i % 2048 + 1) + b"stZ2023-01-0%dT10:00:00Z\0" % (i % 9 + 1) + \
            b"duf" + struct.pack("<f", 1.5 + i)
        rec = synth.bam_record(b"read%d" % i, seq, qual, tags, flag=int(flags[i % len(flags)]))
        if missing_every and i % missing_every == 0 and ln:
            # qualities absent: the block is 0xff bytes (SAM spec 4.2.3; _qcmodule.c:1658-1663)
            body = bytearray(rec)
            qoff = 4 + 32 + len(b"read%d" % i) + 1 + (ln + 1) // 2
            body[qoff:qoff + ln] = b"\xff" * ln
            rec = bytes(body)
        out.write(rec)
    return out.getvalue()


def records(mod, raw, bufsize=None):
    """Every record a BamParser yields, as plain tuples."""
    parser = mod.BamParser(io.BytesIO(raw)) if bufsize is None else mod.BamParser(io.BytesIO(raw), bufsize)
    out = []
    for arr in parser:
        for i in range(len(arr)):
            r = arr[i]
            out.append((r.name(), r.sequence(), r.qualities(), r.tags()))
    return parser.header, out


def outcome(fn):
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        return ("raised", type(e).__name__)


def test_missing_qualities_and_skipped_flags(sq):
    rng = np.random.default_rng(31)
    raw = make_bam(rng, 120, flags=(4, 0x100, 4, 0x800, 0x904, 77, 141), missing_every=3)
    want, got = records(REF, raw), records(sq, raw)
    assert got == want
    assert any(set(q) == {"!"} for _, s, q, _ in want[1] if s)      # the 0xff branch was taken
    assert len(want[1]) < 120                                     # records were skipped
    H.assert_same(H.api_single_end(sq, b"", H.NANOPORE_ADAPTERS, fileobj=io.BytesIO(raw), bam=True),
                  H.api_single_end(REF, b"", H.NANOPORE_ADAPTERS, fileobj=io.BytesIO(raw), bam=True))


def test_only_skipped_records(sq):
    raw = make_bam(np.random.default_rng(32), 9, flags=(0x100, 0x800))
    assert records(sq, raw) == records(REF, raw) and records(REF, raw)[1] == []


@pytest.mark.parametrize("bufsize", [4, 8, 10, 20, 40, 333, 4096])
def test_small_read_steps(sq, bufsize):
    raw = make_bam(np.random.default_rng(33), 40, flags=(4, 4, 0x100), missing_every=5)
    assert records(sq, raw, bufsize) == records(REF, raw, bufsize)


def test_every_truncation_point(sq):
    """tests/test_bam_parser.py:95-120: cut the stream anywhere in the header or in a record."""
    raw = make_bam(np.random.default_rng(34), 3, lengths=(7, 8, 5))
    header_len = len(synth.bam_header())
    for end in range(len(raw)):
        a = outcome(lambda: records(REF, raw[:end]))
        b = outcome(lambda: records(sq, raw[:end]))
        assert a == b, (end, header_len, a, b)
    with pytest.raises(EOFError, match="ncomplete record"):
        records(sq, raw[:header_len + 9])
    with pytest.raises(EOFError, match="runcated BAM"):
        records(sq, raw[:header_len - 3])


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference fixtures not staged")
@pytest.mark.parametrize("name", ["missing_quals.bam", "test_skip.bam", "secondary_alignment.bam",
                                  "simple.unaligned.bam", "simple.raw.bam", "dorado_nanopore_100reads.bam",
                                  "project.NIST_NIST7035_H7AP8ADXX_TAAGGCGA_1_NA12878.bwa.markDuplicates.bam"])
def test_reference_fixture_files(sq, name):
    with open(os.path.join(REF_DATA, name), "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    assert records(sq, raw) == records(REF, raw)
    H.assert_same(H.api_single_end(sq, b"", H.NANOPORE_ADAPTERS, fileobj=io.BytesIO(raw), bam=True),
                  H.api_single_end(REF, b"", H.NANOPORE_ADAPTERS, fileobj=io.BytesIO(raw), bam=True))


# ---- the record chain on the device (sq_bam_walk_device) against the host walk (sq_bam_walk) ----
def _walks(raw, n_ref=0):
    """(kept offsets, n_skipped, consumed) of the host walk and of the device walk over the same bytes."""
    import ctypes as C
    from sequali_b200 import _lib
    ctx = _lib.Context.get()
    out = []
    for device in (False, True):
        offs = np.zeros(len(raw) // 36 + 2, dtype=np.uint64)
        kept, skipped, used = C.c_uint64(), C.c_uint64(), C.c_uint64()
        buf = np.frombuffer(raw, dtype=np.uint8)
        if device:
            rc = ctx.lib.sq_bam_walk_device(ctx.h, buf.ctypes.data, len(raw), n_ref, offs.ctypes.data, len(offs),
                                            C.byref(kept), C.byref(skipped), C.byref(used))
        else:
            rc = ctx.lib.sq_bam_walk(buf.ctypes.data, len(raw), offs.ctypes.data, len(offs), C.byref(kept),
                                     C.byref(skipped), C.byref(used))
        out.append((rc, offs[:kept.value].tolist(), skipped.value, used.value))
    return out


def _body(raw):
    return raw[len(synth.bam_header()):]


@pytest.mark.parametrize("cut", [0, 1, 3, 4, 5, 35, 36, 40, 100, -1, -5, -40])
def test_device_walk_equals_host_walk_at_every_kind_of_end(cut):
    rng = np.random.default_rng(41)
    body = _body(make_bam(rng, 300, flags=(4, 0x104, 4, 0x804, 4, 4)))
    raw = body if cut == 0 else body[:cut]
    host, dev = _walks(raw)
    assert host[0] == 0 and dev == host


def test_device_walk_ignores_headers_that_only_look_like_records():
    """A whole, self-consistent record header inside the tags of a record is a candidate the chain never reaches."""
    rng = np.random.default_rng(42)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    inner = synth.bam_record(b"decoy", letters[rng.integers(0, 4, 40)], np.full(40, 53, np.uint8), b"")
    recs = []
    for i in range(200):
        seq = letters[rng.integers(0, 4, 60 + i)]
        tags = b"xxZ" + inner * (1 + i % 3) + b"\0" if i % 2 else b"chS\x01\x00"
        recs.append(synth.bam_record(b"r%d" % i, seq, np.full(len(seq), 60, np.uint8), tags, flag=4 if i % 7 else 0x904))
    raw = b"".join(recs)
    host, dev = _walks(raw)
    assert host[0] == 0 and len(host[1]) + host[2] == 200 and dev == host


def test_device_walk_follows_records_whose_header_is_not_self_consistent():
    """The reference only trusts block_size: records with an odd reference id, no NUL behind the name or no name
    at all are part of the chain; the device falls back to the plain walk from the first such record."""
    rng = np.random.default_rng(43)
    body = bytearray(_body(make_bam(rng, 120, lengths=(50, 151, 7))))
    host0, _ = _walks(bytes(body))
    offs = host0[1]
    r = offs[40]
    body[r + 4:r + 8] = struct.pack("<i", 77)          # refID beyond n_ref = 0
    r = offs[80]
    name_end = r + 36 + body[r + 12] - 1
    body[name_end] = ord("x")                          # name not NUL terminated
    host, dev = _walks(bytes(body))
    assert host[0] == 0 and host[1] == offs and dev == host
    # with a reference dictionary that covers the id, record 40 is an ordinary candidate again
    host, dev = _walks(bytes(body[:offs[80]]), n_ref=100)
    assert dev == host and len(dev[1]) == 80


def test_device_walk_of_a_chain_that_starts_with_an_odd_record():
    rng = np.random.default_rng(44)
    body = bytearray(_body(make_bam(rng, 50)))
    body[4:8] = struct.pack("<i", 5)
    host, dev = _walks(bytes(body))
    assert host[0] == 0 and len(host[1]) == 50 and dev == host


def test_device_walk_large(sq):
    raw = _body(synth.nanopore_ubam(400, mean_length=3000, max_length=60_000, seed=9))
    host, dev = _walks(raw + raw[:1000])
    assert host[0] == 0 and len(host[1]) == 400 and dev == host


def test_long_records_decode_in_tiles(sq):
    """Reads of 0 .. 3 tiles and every alignment of the output words, against the unmodified reference."""
    rng = np.random.default_rng(45)
    lengths = (0, 1, 4095, 4096, 4097, 8191, 8192, 8193, 12289, 3, 20000, 5)
    raw = make_bam(rng, 36, flags=(4,), missing_every=5, lengths=lengths)
    want = records(REF, raw)
    assert records(sq, raw) == want
    assert records(sq, raw, 1000) == want
