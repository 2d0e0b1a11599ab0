"""-m gpu: the single-record API (add_read / add_sequence[_pair], members read right after a call,
exceptions, constants) run side by side on the unmodified reference extension (oracle/_ref, built by
oracle/build_ref.sh) and on sequali_b200 -- same calls, same answers.  These are the call patterns of
the reference's own tests (SURVEY.md 8b "test-pinned behaviours")."""
import warnings

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

REF = H.import_reference()
if REF is None:  # pragma: no cover
    pytest.skip("oracle/_ref is not built", allow_module_level=True)


@pytest.fixture(scope="module", params=["ctypes", "extension"])
def sq(request):
    """The B200 build behind its two host layers: the ctypes mirror and the CPython extension."""
    if request.param == "extension":
        import sequali_b200.ext
        return sequali_b200.ext
    import sequali_b200
    return sequali_b200


def outcome(fn):
    """Result of fn(), or (exception type, message); warnings are returned too."""
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            res = fn()
        except Exception as e:  # noqa: BLE001
            res = ("raised", type(e).__name__, str(e))
    return res, [(x.category.__name__, str(x.message)) for x in w]


def both(sq, script):
    a, b = outcome(lambda: script(REF)), outcome(lambda: script(sq))
    assert a == b
    return a


def views(mod, rng, n, tile=True, lengths=(0, 1, 5, 37, 150, 151, 400)):
    out = []
    for i in range(n):
        L = int(lengths[i % len(lengths)])
        seq = "".join(rng.choice(list("ACGTacgtN"), L, p=[.22, .22, .22, .22, .02, .02, .02, .02, .04]))
        qual = "".join(chr(33 + int(q)) for q in rng.integers(0, 94, L))
        name = f"SIM:1:FCX:1:{1101 + i % 3}:{i}:{i * 7} 1:N:0:ATCACG" if tile else f"read{i}"
        out.append(mod.FastqRecordView(name, seq, qual))
    return out


def test_record_views_and_arrays(sq):
    def script(m):
        v = m.FastqRecordView("name more", "ACGTN", "IIII!", b"RGZx\x00")
        arr = m.FastqRecordArrayView([v, m.FastqRecordView("n2", "", "")])
        pair = m.FastqRecordArrayView([m.FastqRecordView("name 2:N", "A", "I"), m.FastqRecordView("n2/2", "", "")])
        return (v.name(), v.sequence(), v.qualities(), v.tags(), len(arr), arr[0].name(), arr[1].sequence(),
                arr[-1].name(), arr.is_mate(pair), pair.is_mate(arr), isinstance(arr.obj, bytes))
    both(sq, script)
    for bad in (lambda m: m.FastqRecordView("n", "ACGT", "III"), lambda m: m.FastqRecordView("n", "A", " "),
                lambda m: m.FastqRecordArrayView([m.FastqRecordView("n", "A", "I")])[3],
                lambda m: m.FastqRecordArrayView([m.FastqRecordView("n", "A", "I")]).is_mate("nope")):
        res, _ = both(sq, bad)
        assert res[0] == "raised"
    # (the reference's own message for these two is a broken format string: only "it raises" is compared)
    for bad in (lambda m: m.FastqRecordArrayView([1, 2]), lambda m: m.FastqRecordArrayView("x")):
        assert outcome(lambda: bad(REF))[0][0] == "raised" and outcome(lambda: bad(sq))[0][0] == "raised"


def test_qc_metrics_add_read(sq):
    def script(m):
        rng = np.random.default_rng(1)
        q = m.QCMetrics(end_anchor_length=40)
        snap = []
        for v in views(m, rng, 40):
            q.add_read(v)
            snap.append((q.number_of_reads, q.max_length))
        return (snap, q.end_anchor_length, q.base_count_table().tolist(), q.phred_count_table().tolist(),
                q.end_anchored_base_count_table().tolist(), q.end_anchored_phred_count_table().tolist(),
                q.gc_content().tolist(), q.phred_scores().tolist())
    both(sq, script)
    res, _ = both(sq, lambda m: m.QCMetrics().add_read("not a view"))
    assert res[0] == "raised" and res[1] == "TypeError"
    res, _ = both(sq, lambda m: m.QCMetrics().add_record_array([1]))
    assert res[0] == "raised" and res[1] == "TypeError"


def test_adapter_counter_add_read(sq):
    adapters = ["AGATCGGAAGAG", "ACGT", "GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG", "TTN"]

    def script(m):
        rng = np.random.default_rng(2)
        a = m.AdapterCounter(adapters)
        for v in views(m, rng, 60):
            a.add_read(v)
        a.add_read(m.FastqRecordView("x", "TTAGATCGGAAGAGTTACGT", "I" * 20))
        return (a.number_of_sequences, a.max_length, list(a.adapters),
                [(ad, f.tolist(), r.tolist()) for ad, f, r in a.get_counts()])
    both(sq, script)
    for bad in (lambda m: m.AdapterCounter([]), lambda m: m.AdapterCounter(["AC", 5]),
                lambda m: list(m.AdapterCounter("ACGT").adapters), lambda m: m.AdapterCounter(["A" * 65])):
        both(sq, bad)


def test_per_tile_quality_add_read(sq):
    def script(m):
        rng = np.random.default_rng(3)
        p = m.PerTileQuality()
        snap = []
        for v in views(m, rng, 45):
            p.add_read(v)
            snap.append((p.number_of_reads, p.max_length))
        before = p.skipped_reason
        p.add_read(m.FastqRecordView("no tile here", "ACGT", "IIII"))
        p.add_read(m.FastqRecordView("SIM:1:FCX:1:1101:5:5", "ACGT", "IIII"))
        return snap, before, p.skipped_reason, p.number_of_reads, p.get_tile_counts()
    both(sq, script)


def test_overrepresented_sequences_add_read(sq):
    def script(m):
        rng = np.random.default_rng(4)
        o = m.OverrepresentedSequences(max_unique_fragments=50, fragment_length=7, sample_every=2,
                                       bases_from_start=30, bases_from_end=30)
        for v in views(m, rng, 80, lengths=(0, 6, 7, 8, 30, 61, 150)):
            o.add_read(v)
        o.add_read(m.FastqRecordView("odd", "ACGTKKKACGTACGT", "I" * 15))
        o.add_read(m.FastqRecordView("odd2", "ACGTKKKACGTACGT", "I" * 15))
        return (o.number_of_sequences, o.sampled_sequences, o.collected_unique_fragments, o.max_unique_fragments,
                o.fragment_length, o.sample_every, o.total_fragments, sorted(o.sequence_counts().items()),
                o.overrepresented_sequences(threshold_fraction=0.01, min_threshold=1, max_threshold=5))
    both(sq, script)
    for bad in (lambda m: m.OverrepresentedSequences(fragment_length=8), lambda m: m.OverrepresentedSequences(0),
                lambda m: m.OverrepresentedSequences(sample_every=0)):
        res, _ = both(sq, bad)
        assert res[0] == "raised"


def test_dedup_estimator_add_sequence(sq):
    def script(m):
        rng = np.random.default_rng(5)
        d = m.DedupEstimator(max_stored_fingerprints=100, front_sequence_length=4, back_sequence_length=4,
                             front_sequence_offset=2, back_sequence_offset=2)
        seqs = ["".join(rng.choice(list("ACGT"), int(L))) for L in rng.integers(0, 40, 900)]
        for s in seqs + seqs[:50]:
            d.add_sequence(s)
        for s1, s2 in zip(seqs[:40], seqs[40:80]):
            d.add_sequence_pair(s1, s2)
        return d._modulo_bits, d._hash_table_size, d.tracked_sequences, sorted(d.duplication_counts().tolist())
    both(sq, script)
    res, _ = both(sq, lambda m: m.DedupEstimator().add_sequence(b"ACGT"))
    assert res[0] == "raised"
    both(sq, lambda m: m.DedupEstimator(max_stored_fingerprints=0))


def test_nano_stats_add_read(sq):
    def script(m):
        rng = np.random.default_rng(6)
        n = m.NanoStats()
        for i in range(30):
            L = int(rng.integers(0, 300))
            name = (f"{'%08x' % i}-0000-4000-8000-{'%012x' % (i * 977)} runid=ab{i} read={i} ch={1 + i % 512} "
                    f"start_time=2023-05-{1 + i % 28:02d}T1{i % 10}:0{i % 6}:1{i % 10}Z")
            n.add_read(m.FastqRecordView(name, "A" * L, "".join(chr(33 + int(q)) for q in rng.integers(0, 60, L))))
        infos = [(x.start_time, x.channel_id, x.length, x.cumulative_error_rate, x.duration, x.parent_id_hash)
                 for x in n.nano_info_iterator()]
        before = n.skipped_reason
        n.add_read(m.FastqRecordView("not nanopore", "ACGT", "IIII"))
        return infos, n.number_of_reads, n.minimum_time, n.maximum_time, before, n.skipped_reason
    both(sq, script)


def test_insert_size_metrics_add_sequence_pair(sq):
    comp = str.maketrans("ACGT", "TGCA")

    def script(m):
        rng = np.random.default_rng(7)
        ins = m.InsertSizeMetrics(max_adapters=8)
        a1, a2 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA", "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
        for i in range(120):
            size = int(rng.integers(20, 260))
            frag = "".join(rng.choice(list("ACGT"), size))
            r1 = (frag + a1 + "A" * 100)[:100]
            r2 = (frag[::-1].translate(comp) + a2 + "A" * 100)[:100]
            ins.add_sequence_pair(r1, r2)
        return (ins.total_reads, ins.number_of_adapters_read1, ins.number_of_adapters_read2,
                ins.insert_sizes().tolist(), ins.adapters_read1(), ins.adapters_read2())
    both(sq, script)


def test_module_constants(sq):
    names = [n for n in dir(REF) if n.isupper()]
    assert len(names) > 10
    for n in names:
        assert getattr(sq, n) == getattr(REF, n), n


# ----------------------------------------------------------------------------
# whole runs, reference extension against CUDA path directly (no oracle in between)
# ----------------------------------------------------------------------------
from sequali_b200 import synth  # noqa: E402


@pytest.mark.parametrize("bufsize", [128 * 1024, 8 << 20])
def test_single_end_run_equals_the_reference(sq, bufsize):
    text = synth.illumina_fastq(60_000, length=150, seed=77, n_tiles=60)
    kw = dict(dedup_kwargs=dict(max_stored_fingerprints=3000), overrep_kwargs=dict(max_unique_fragments=20_000))
    want = H.api_single_end(REF, text, H.ILLUMINA_ADAPTERS, buffersize=128 * 1024, **kw)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=bufsize, **kw)
    H.assert_same(got, want)


def test_variable_length_random_tiles_equals_the_reference(sq):
    text = synth.illumina_fastq(25_000, length=120, seed=78, n_tiles=150, tile_runs=False, variable_length=True)
    want = H.api_single_end(REF, text, H.ILLUMINA_ADAPTERS)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=1 << 20)
    H.assert_same(got, want)


def test_nanopore_run_equals_the_reference(sq):
    text = synth.nanopore_fastq(300, mean_length=8000, max_length=200_000, seed=79)
    want = H.api_single_end(REF, text, H.NANOPORE_ADAPTERS)
    got = H.api_single_end(sq, text, H.NANOPORE_ADAPTERS, buffersize=4 << 20)
    H.assert_same(got, want)


def test_paired_run_equals_the_reference(sq):
    t1, t2 = synth.paired_fastq(30_000, seed=80)
    H.assert_same(H.api_paired(sq, t1, t2, buffersize=2 << 20), H.api_paired(REF, t1, t2))
