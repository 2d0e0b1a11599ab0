"""Seeded synthetic inputs of the shapes BASELINE.json names.

Host-side (numpy) generators used by the parity tests and by bench.py for the
CPU-sized samples.  The header style follows the reference's fixture generator
(scripts/fastq_create.py:8): ``SIM:1:FCX:1:<tile>:<x>:<y> 1:N:0:ATCACG``.
The 100 M-read bench input is produced on the device by the same recipe
(csrc/synth_kernels.cu) so that it never has to cross PCIe.
"""
from __future__ import annotations

import io
import struct
import uuid

import numpy as np

ILLUMINA_ADAPTER_R1 = b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
ILLUMINA_ADAPTER_R2 = b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def novaseq_tiles() -> list[int]:
    """{1,2}{1..6}{01..78}: 936 tiles in flow-cell order."""
    return [s * 1000 + w * 100 + t for s in (1, 2) for w in range(1, 7) for t in range(1, 79)]


def _random_bases(rng, n, length, n_frac=0.001):
    seq = _ACGT[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]
    if n_frac:
        seq[rng.random((n, length)) < n_frac] = ord("N")
    return seq


def _illumina_quals(rng, n, length):
    mean = np.clip(rng.normal(34, 4, size=(n, 1)), 2, 41)
    decay = np.linspace(0, 6, length)[None, :] * rng.random((n, 1))
    q = mean - rng.integers(0, 9, size=(n, length)) - decay
    return (np.clip(np.rint(q), 2, 41).astype(np.uint8) + 33)


def _illumina_names(rng, n, tiles, runs=True, mate=1):
    if runs:
        tile_idx = np.sort(rng.integers(0, len(tiles), size=n))
    else:
        tile_idx = rng.integers(0, len(tiles), size=n)
    x = rng.integers(1000, 32000, size=n)
    y = rng.integers(1000, 200000, size=n)
    return [b"SIM:1:FCX:1:%d:%d:%d %d:N:0:ATCACG" % (tiles[t], a, b, mate)
            for t, a, b in zip(tile_idx.tolist(), x.tolist(), y.tolist())]


def _assemble(names, seq, qual, lengths=None) -> bytes:
    out = io.BytesIO()
    if lengths is None:
        for nm, s, q in zip(names, seq, qual):
            out.write(b"@" + nm + b"\n" + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    else:
        for nm, s, q, ln in zip(names, seq, qual, lengths):
            out.write(b"@" + nm + b"\n" + s[:ln].tobytes() + b"\n+\n" + q[:ln].tobytes() + b"\n")
    return out.getvalue()


def illumina_fastq(n_reads: int, length: int = 150, seed: int = 1, n_tiles: int = 192,
                   tile_runs: bool = True, adapter_frac: float = 0.05,
                   dup_frac: float = 0.02, variable_length: bool = False) -> bytes:
    """C1/C2-style single-end Illumina reads."""
    rng = np.random.default_rng(seed)
    seq = _random_bases(rng, n_reads, length)
    qual = _illumina_quals(rng, n_reads, length)
    for i in np.flatnonzero(rng.random(n_reads) < adapter_frac):
        p = int(rng.integers(20, length - 12))
        tail = ILLUMINA_ADAPTER_R1 + _ACGT[rng.integers(0, 4, size=length)].tobytes()
        seq[i, p:] = np.frombuffer(tail[:length - p], dtype=np.uint8)
    dups = np.flatnonzero(rng.random(n_reads) < dup_frac)
    dups = dups[dups > 0]
    if dups.size:
        seq[dups] = seq[(rng.random(dups.size) * dups).astype(np.int64)]
    tiles = novaseq_tiles()[:n_tiles]
    names = _illumina_names(rng, n_reads, tiles, runs=tile_runs)
    lengths = None
    if variable_length:
        lengths = rng.integers(0, length + 1, size=n_reads)
    return _assemble(names, seq, qual, lengths)


def paired_fastq(n_pairs: int, length: int = 150, seed: int = 3, n_tiles: int = 96,
                 error_rate: float = 0.01) -> tuple[bytes, bytes]:
    """C3-style pairs: insert ~ N(220,60) in [30,600], R2 = revcomp of the insert
    end, adapter read-through when insert < length, substitution errors."""
    rng = np.random.default_rng(seed)
    insert = np.clip(np.rint(rng.normal(220, 60, size=n_pairs)), 30, 600).astype(np.int64)
    frag = _random_bases(rng, n_pairs, 600, n_frac=0.0)
    r1 = np.empty((n_pairs, length), dtype=np.uint8)
    r2 = np.empty((n_pairs, length), dtype=np.uint8)
    pad = _ACGT[rng.integers(0, 4, size=(n_pairs, length), dtype=np.uint8)]
    a1 = np.frombuffer(ILLUMINA_ADAPTER_R1, dtype=np.uint8)
    a2 = np.frombuffer(ILLUMINA_ADAPTER_R2, dtype=np.uint8)
    for i in range(n_pairs):
        ins = int(insert[i])
        f = frag[i, :ins]
        rc = _COMP[f[::-1]]
        if ins >= length:
            r1[i] = f[:length]
            r2[i] = rc[:length]
        else:
            t1 = np.concatenate([f, a1, pad[i]])[:length]
            t2 = np.concatenate([rc, a2, pad[i]])[:length]
            r1[i], r2[i] = t1, t2
    for r in (r1, r2):
        mask = rng.random(r.shape) < error_rate
        r[mask] = _ACGT[rng.integers(0, 4, size=int(mask.sum()), dtype=np.uint8)]
    q1 = _illumina_quals(rng, n_pairs, length)
    q2 = _illumina_quals(rng, n_pairs, length)
    tiles = novaseq_tiles()[:n_tiles]
    names1 = _illumina_names(rng, n_pairs, tiles, mate=1)
    names2 = [nm.replace(b" 1:", b" 2:") for nm in names1]
    return _assemble(names1, r1, q1), _assemble(names2, r2, q2)


NANOPORE_PROBES = [
    b"TTACGTATTGCT", b"GCAATACGTAAC", b"CTTGCGGGCGGC", b"GGTAGTAGGTTC", b"GAGGCGAGCGGT",
    b"CAAGATACGCAC", b"GTGACTTGCCTG", b"ATCGCCTACCGT", b"TCTATCTTCTTT", b"TCTTCAGAGGAG",
    b"GATATTGCTGGG", b"TGATATTGCTTT", b"GTACGTATTGCT", b"ACGTAACTGAAC",
]


def _nanopore_reads(rng, n_reads, mean_length, max_length):
    lengths = np.minimum(rng.gamma(1.2, mean_length / 1.2, size=n_reads).astype(np.int64) + 200,
                         max_length)
    reads = []
    for ln in lengths.tolist():
        s = _ACGT[rng.integers(0, 4, size=ln, dtype=np.uint8)]
        if rng.random() < 0.10:
            probe = np.frombuffer(NANOPORE_PROBES[int(rng.integers(0, len(NANOPORE_PROBES)))],
                                  dtype=np.uint8)
            at = int(rng.integers(0, 88)) if rng.random() < 0.5 else ln - 12 - int(rng.integers(0, 88))
            s[at:at + 12] = probe
        q = rng.integers(3, 46, size=ln, dtype=np.uint8) + 33
        reads.append((s, q))
    return reads


def _rand_uuid(rng) -> str:
    return str(uuid.UUID(bytes=rng.bytes(16), version=4))


def _iso_time(rng, span_hours=48):
    t = 1632000000 + int(rng.integers(0, span_hours * 3600))
    d = np.datetime64(t, "s")
    return str(d) + "Z", t


def nanopore_fastq(n_reads: int, mean_length: int = 20000, max_length: int = 1_000_000,
                   seed: int = 4) -> bytes:
    """C4-style reads with guppy headers (ch= and start_time= fields)."""
    rng = np.random.default_rng(seed)
    out = io.BytesIO()
    runid = rng.bytes(20).hex()
    for i, (s, q) in enumerate(_nanopore_reads(rng, n_reads, mean_length, max_length)):
        ts, _ = _iso_time(rng)
        name = "%s runid=%s read=%d ch=%d start_time=%s" % (
            _rand_uuid(rng), runid, i, int(rng.integers(1, 2049)), ts)
        out.write(b"@" + name.encode() + b"\n" + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return out.getvalue()


_NIB = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"=ACMGRSVTWYHKDBN"):
    _NIB[_c] = _i


def bam_record(name: bytes, seq: np.ndarray, qual: np.ndarray, tags: bytes, flag: int = 4) -> bytes:
    """One unaligned BAM alignment record (SAM spec §4.2); qual is phred+33."""
    ln = len(seq)
    nib = _NIB[seq]
    if ln & 1:
        nib = np.concatenate([nib, np.zeros(1, dtype=np.uint8)])
    packed = (nib[0::2] << 4 | nib[1::2]).astype(np.uint8).tobytes()
    body = struct.pack("<iiBBHHHIiii", -1, -1, len(name) + 1, 0, 4680, 0, flag, ln, -1, -1, 0)
    body += name + b"\0" + packed + (qual - 33).astype(np.uint8).tobytes() + tags
    return struct.pack("<I", len(body)) + body


def bam_header(text: bytes = b"@HD\tVN:1.6\tSO:unknown\n@RG\tID:rg1\tPL:ONT\n") -> bytes:
    return b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 0)


def nanopore_ubam(n_reads: int, mean_length: int = 20000, max_length: int = 1_000_000,
                  seed: int = 5) -> bytes:
    """C5: dorado-style unaligned BAM (uncompressed stream, header included)."""
    rng = np.random.default_rng(seed)
    out = io.BytesIO()
    out.write(bam_header())
    for i, (s, q) in enumerate(_nanopore_reads(rng, n_reads, mean_length, max_length)):
        ts, _ = _iso_time(rng)
        tags = b"qsC" + bytes([int(rng.integers(5, 40))])
        tags += b"duf" + struct.pack("<f", float(rng.random() * 30))
        tags += b"nsS" + struct.pack("<H", int(rng.integers(0, 65535)))
        tags += b"tsC" + bytes([int(rng.integers(0, 200))])
        tags += b"mxC" + bytes([1])
        tags += b"chS" + struct.pack("<H", int(rng.integers(1, 2049)))
        tags += b"stZ" + ts.encode() + b"\0"
        tags += b"rnI" + struct.pack("<I", i)
        tags += b"fnZ" + b"file_%d.pod5\0" % (i % 7)
        tags += b"smf" + struct.pack("<f", 90.5) + b"sdf" + struct.pack("<f", 17.25)
        tags += b"svZ" + b"quantile\0" + b"dxC" + bytes([0]) + b"RGZ" + b"rg1\0"
        if rng.random() < 0.05:
            tags += b"piZ" + _rand_uuid(rng).encode() + b"\0"
        out.write(bam_record(_rand_uuid(rng).encode(), s, q, tags))
    return out.getvalue()


def bgzf_compress(data: bytes, level: int = 6, block_text: int = 65280, eof_marker: bool = True) -> bytes:
    """`data` as a BGZF stream (SAM spec 4.1: what bgzip writes): independent gzip members of
    `block_text` bytes of text each, every one with the 'BC' extra field giving its size."""
    import zlib
    out = io.BytesIO()

    def member(chunk: bytes):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        total = 12 + 6 + len(payload) + 8
        assert total <= 65536, "pick a smaller block_text for incompressible data"
        out.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) +
                  b"BC" + struct.pack("<HH", 2, total - 1) + payload +
                  struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    for i in range(0, len(data), block_text):
        member(data[i:i + block_text])
    if eof_marker:
        member(b"")
    return out.getvalue()
