"""Shared helpers for the parity tests: run the same input through the real
reference (oracle/_ref, when present), the C restatement (oracle/) and, in the
``-m gpu`` tests, the CUDA build, and reduce every getter to a canonical,
directly comparable form (doubles by bit pattern)."""
from __future__ import annotations

import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "data")

from oracle import oracle as orc  # noqa: E402

ILLUMINA_ADAPTERS = ["AGATCGGAAGAG", "TGGAATTCTCGG", "GATCGTCGGACT", "CTGTCTCTTATA",
                     "GGGGGGGGGGGG", "AAAAAAAAAAAA"]
NANOPORE_ADAPTERS = ["TTACGTATTGCT", "GCAATACGTAAC", "CTTGCGGGCGGC", "GGTAGTAGGTTC",
                     "GAGGCGAGCGGT", "CAAGATACGCAC", "GTGACTTGCCTG", "ATCGCCTACCGT",
                     "TCTATCTTCTTT", "TCTTCAGAGGAG", "GATATTGCTGGG", "TGATATTGCTTT",
                     "GTACGTATTGCT", "ACGTAACTGAAC"]


def import_reference():
    """The unmodified reference extension built by oracle/build_ref.sh, or None."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "sequali", "_qc.abi3.so")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import sequali  # type: ignore
        return sequali
    except Exception:
        return None


def f64_bits(a) -> list[int]:
    return np.asarray(a, dtype="<f8").view("<u8").tolist()


# ----------------------------------------------------------------------------
# canonical dumps of API-shaped objects (the reference, or sequali_b200)
# ----------------------------------------------------------------------------

def dump_qc(m) -> dict:
    return dict(
        max_length=m.max_length, number_of_reads=m.number_of_reads,
        base=list(m.base_count_table()), phred=list(m.phred_count_table()),
        ea_base=list(m.end_anchored_base_count_table()),
        ea_phred=list(m.end_anchored_phred_count_table()),
        gc=list(m.gc_content()), mean_phred=list(m.phred_scores()))


def dump_adapters(a) -> dict:
    return dict(max_length=a.max_length, number_of_sequences=a.number_of_sequences,
                counts=[(ad, list(f), list(r)) for ad, f, r in a.get_counts()])


def dump_ptq(p) -> dict:
    return dict(max_length=p.max_length, number_of_reads=p.number_of_reads,
                skipped_reason=p.skipped_reason,
                tiles=[(t, f64_bits(e), list(c)) for t, e, c in p.get_tile_counts()])


def dump_overrep(o) -> dict:
    return dict(number_of_sequences=o.number_of_sequences, sampled_sequences=o.sampled_sequences,
                collected_unique_fragments=o.collected_unique_fragments,
                total_fragments=o.total_fragments, counts=dict(o.sequence_counts()))


def dump_dedup(d) -> dict:
    counts = list(d.duplication_counts())
    return dict(modulo_bits=d._modulo_bits, tracked_sequences=d.tracked_sequences,
                table_size=d._hash_table_size, slot_order=counts, sorted=sorted(counts))


def dump_nano(n) -> dict:
    infos = [(i.start_time, i.channel_id, i.length, f64_bits([i.cumulative_error_rate])[0],
              np.float32(i.duration).view(np.uint32).item(), i.parent_id_hash)
             for i in n.nano_info_iterator()]
    return dict(number_of_reads=n.number_of_reads, minimum_time=n.minimum_time,
                maximum_time=n.maximum_time, skipped_reason=n.skipped_reason, infos=infos)


def dump_insert(m) -> dict:
    a1, a2 = list(m.adapters_read1()), list(m.adapters_read2())
    return dict(total_reads=m.total_reads, number_of_adapters_read1=m.number_of_adapters_read1,
                number_of_adapters_read2=m.number_of_adapters_read2,
                sizes=list(m.insert_sizes()), adapters1=sorted(a1), adapters2=sorted(a2),
                adapters1_slot_order=a1, adapters2_slot_order=a2)


# ----------------------------------------------------------------------------
# canonical dumps of the oracle objects (same keys)
# ----------------------------------------------------------------------------

def odump_qc(m: orc.QCMetrics) -> dict:
    t = m.tables()
    return dict(max_length=m.max_length, number_of_reads=m.number_of_reads,
                **{k: t[k].tolist() for k in ("base", "phred", "ea_base", "ea_phred", "gc",
                                               "mean_phred")})


def odump_adapters(a: orc.AdapterCounter) -> dict:
    return dict(max_length=a.max_length, number_of_sequences=a.number_of_sequences,
                counts=[(ad, f.tolist(), r.tolist()) for ad, f, r in a.get_counts()])


def odump_ptq(p: orc.PerTileQuality, skipped_reason=None) -> dict:
    return dict(max_length=p.max_length, number_of_reads=p.number_of_reads,
                skipped_reason=skipped_reason,
                tiles=[(t, f64_bits(e), c.tolist()) for t, e, c in p.get_tile_counts()])


def odump_overrep(o: orc.OverrepresentedSequences) -> dict:
    i = o.info()
    return dict(number_of_sequences=i["number_of_sequences"],
                sampled_sequences=i["sampled_sequences"],
                collected_unique_fragments=i["collected_unique_fragments"],
                total_fragments=i["total_fragments"], counts=o.sequence_counts())


def odump_dedup(d: orc.DedupEstimator) -> dict:
    i = d.info()
    counts = d.duplication_counts().tolist()
    return dict(modulo_bits=i["_modulo_bits"], tracked_sequences=i["tracked_sequences"],
                table_size=i["_hash_table_size"], slot_order=counts, sorted=sorted(counts))


def odump_nano(n: orc.NanoStats, skipped_reason=None) -> dict:
    i = n.info()
    infos = [(int(r["start_time"]), int(r["channel_id"]), int(r["length"]),
              f64_bits([r["cumulative_error_rate"]])[0],
              np.float32(r["duration"]).view(np.uint32).item(), int(r["parent_id_hash"]))
             for r in n.infos()]
    return dict(number_of_reads=i["number_of_reads"], minimum_time=i["minimum_time"],
                maximum_time=i["maximum_time"], skipped_reason=skipped_reason, infos=infos)


def odump_insert(m: orc.InsertSizeMetrics) -> dict:
    i = m.info()
    a1, a2 = m.adapters(1), m.adapters(2)
    return dict(total_reads=i["total_reads"],
                number_of_adapters_read1=i["number_of_adapters_read1"],
                number_of_adapters_read2=i["number_of_adapters_read2"],
                sizes=m.insert_sizes().tolist(), adapters1=sorted(a1), adapters2=sorted(a2),
                adapters1_slot_order=a1, adapters2_slot_order=a2)


# ----------------------------------------------------------------------------
# whole-pipeline runs, shaped like src/sequali/__main__.py:214-306
# ----------------------------------------------------------------------------

def _header_repr(buf, rec) -> str:
    name = bytes(buf[int(rec["name_off"]):int(rec["name_off"]) + int(rec["name_len"])])
    return "Can not parse header: %r" % name.decode("ascii")


def oracle_single_end(text: bytes, adapters, chunk_records: int | None = None,
                      dedup_kwargs=None, overrep_kwargs=None) -> dict:
    """All single-end modules through the C restatement.  chunk_records splits the
    record stream into arrays (results must not depend on it)."""
    recs, consumed = orc.parse_fastq(text)
    assert consumed == len(text)
    buf = np.frombuffer(text, dtype=np.uint8)
    qc, ad, ptq = orc.QCMetrics(), orc.AdapterCounter(adapters), orc.PerTileQuality()
    ov = orc.OverrepresentedSequences(**(overrep_kwargs or {}))
    dd = orc.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=64,
                                                     back_sequence_offset=0)))
    ns = orc.NanoStats()
    step = chunk_records or max(len(recs), 1)
    ptq_reason = ns_reason = None
    for s in range(0, len(recs), step):
        part = recs[s:s + step]
        qc.add(buf, part)
        if ptq.add(buf, part) == 1:
            ptq_reason = _header_repr(buf, part[ptq.skipped_record])
        ov.add(buf, part)
        if ns.add(buf, part) == 1:
            ns_reason = _header_repr(buf, part[ns.skipped_record])
        ad.add(buf, part)
        dd.add(buf, part)
    return dict(qc=odump_qc(qc), adapters=odump_adapters(ad), ptq=odump_ptq(ptq, ptq_reason),
                overrep=odump_overrep(ov), dedup=odump_dedup(dd), nano=odump_nano(ns, ns_reason))


def api_single_end(mod, text: bytes, adapters, buffersize: int = 128 * 1024,
                   dedup_kwargs=None, overrep_kwargs=None, fileobj=None, bam=False) -> dict:
    """The same loop through an API-shaped module (the reference or sequali_b200)."""
    qc, ad, ptq = mod.QCMetrics(), mod.AdapterCounter(adapters), mod.PerTileQuality()
    ov = mod.OverrepresentedSequences(**(overrep_kwargs or {}))
    dd = mod.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=64,
                                                     back_sequence_offset=0)))
    ns = mod.NanoStats()
    f = fileobj or io.BytesIO(text)
    parser = mod.BamParser(f, buffersize) if bam else mod.FastqParser(f, buffersize)
    for arr in parser:
        qc.add_record_array(arr)
        ptq.add_record_array(arr)
        ov.add_record_array(arr)
        ns.add_record_array(arr)
        ad.add_record_array(arr)
        dd.add_record_array(arr)
    return dict(qc=dump_qc(qc), adapters=dump_adapters(ad), ptq=dump_ptq(ptq),
                overrep=dump_overrep(ov), dedup=dump_dedup(dd), nano=dump_nano(ns))


def oracle_paired(text1: bytes, text2: bytes, chunk_records=None) -> dict:
    r1, c1 = orc.parse_fastq(text1)
    r2, c2 = orc.parse_fastq(text2)
    assert c1 == len(text1) and c2 == len(text2) and len(r1) == len(r2)
    b1, b2 = np.frombuffer(text1, np.uint8), np.frombuffer(text2, np.uint8)
    qc1, qc2 = orc.QCMetrics(), orc.QCMetrics()
    p1, p2 = orc.PerTileQuality(), orc.PerTileQuality()
    o1, o2 = orc.OverrepresentedSequences(), orc.OverrepresentedSequences()
    dd = orc.DedupEstimator(front_sequence_offset=0, back_sequence_offset=0)
    ins = orc.InsertSizeMetrics()
    step = chunk_records or max(len(r1), 1)
    for s in range(0, len(r1), step):
        a, b = r1[s:s + step], r2[s:s + step]
        qc1.add(b1, a); p1.add(b1, a); o1.add(b1, a)
        dd.add_pair(b1, a, b2, b)
        ins.add_pair(b1, a, b2, b)
        qc2.add(b2, b); p2.add(b2, b); o2.add(b2, b)
    return dict(qc1=odump_qc(qc1), qc2=odump_qc(qc2), ptq1=odump_ptq(p1), ptq2=odump_ptq(p2),
                overrep1=odump_overrep(o1), overrep2=odump_overrep(o2), dedup=odump_dedup(dd),
                insert=odump_insert(ins))


def api_paired(mod, text1: bytes, text2: bytes, buffersize: int = 128 * 1024) -> dict:
    qc1, qc2 = mod.QCMetrics(), mod.QCMetrics()
    p1, p2 = mod.PerTileQuality(), mod.PerTileQuality()
    o1, o2 = mod.OverrepresentedSequences(), mod.OverrepresentedSequences()
    dd = mod.DedupEstimator(front_sequence_offset=0, back_sequence_offset=0)
    ins = mod.InsertSizeMetrics()
    rd1 = mod.FastqParser(io.BytesIO(text1), buffersize)
    rd2 = mod.FastqParser(io.BytesIO(text2), buffersize)
    for a in rd1:
        qc1.add_record_array(a); p1.add_record_array(a); o1.add_record_array(a)
        b = rd2.read(len(a))
        assert len(a) == len(b) and a.is_mate(b)
        dd.add_record_array_pair(a, b)
        ins.add_record_array_pair(a, b)
        qc2.add_record_array(b); p2.add_record_array(b); o2.add_record_array(b)
    assert len(rd2.read(1)) == 0
    return dict(qc1=dump_qc(qc1), qc2=dump_qc(qc2), ptq1=dump_ptq(p1), ptq2=dump_ptq(p2),
                overrep1=dump_overrep(o1), overrep2=dump_overrep(o2), dedup=dump_dedup(dd),
                insert=dump_insert(ins))


def assert_same(a: dict, b: dict, path="", skip=()):
    """Deep equality with a readable first difference."""
    assert a.keys() == b.keys(), (path, sorted(a.keys()), sorted(b.keys()))
    for k in a:
        if k in skip:
            continue
        x, y, p = a[k], b[k], f"{path}.{k}"
        if isinstance(x, dict) and isinstance(y, dict) and k != "counts":
            assert_same(x, y, p, skip)
        elif x != y:
            if isinstance(x, (list, tuple)) and isinstance(y, (list, tuple)):
                assert len(x) == len(y), f"{p}: length {len(x)} != {len(y)}"
                for i, (u, v) in enumerate(zip(x, y)):
                    assert u == v, f"{p}[{i}]: {str(u)[:200]} != {str(v)[:200]}"
            if isinstance(x, dict):
                assert len(x) == len(y), f"{p}: dict sizes {len(x)} != {len(y)}"
                for kk in x:
                    assert kk in y and x[kk] == y[kk], f"{p}[{kk!r}]: {x[kk]} != {y.get(kk)}"
            raise AssertionError(f"{p}: {str(x)[:300]} != {str(y)[:300]}")
