"""-m gpu: the parity gate of SURVEY.md 8(d) at the sizes it names -- C1 in full (1 M x 150 bp
single end, all default modules), C3 / C4 / C5 at 1/100 scale (500 k pairs; 100 Mbases of
ultra-long nanopore reads as FASTQ and as unaligned BAM).  The 6.5 M-read prefix of C2 is
tests/test_gpu_parity.py::test_bench_input_prefix_all_modules.  Every getter of every collector,
bit for bit against the CPU oracle."""
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


def test_c1_one_million_illumina_reads_all_modules(sq):
    text = synth.illumina_fastq(1_000_000, length=150, seed=1, n_tiles=936)
    got = H.api_single_end(sq, text, H.ILLUMINA_ADAPTERS, buffersize=64 << 20)
    want = H.oracle_single_end(text, H.ILLUMINA_ADAPTERS, chunk_records=1 << 16)
    assert want["qc"]["number_of_reads"] == 1_000_000
    H.assert_same(got, want)


def test_c3_paired_end_half_a_million_pairs(sq):
    t1, t2 = synth.paired_fastq(500_000, seed=3)
    got = H.api_paired(sq, t1, t2, buffersize=64 << 20)
    want = H.oracle_paired(t1, t2)
    H.assert_same(got, want)


def test_c4_ultra_long_nanopore_fastq(sq):
    text = synth.nanopore_fastq(5000, mean_length=20_000, max_length=1_000_000, seed=4)
    assert len(text) > 150_000_000
    got = H.api_single_end(sq, text, H.NANOPORE_ADAPTERS, buffersize=32 << 20)
    want = H.oracle_single_end(text, H.NANOPORE_ADAPTERS)
    H.assert_same(got, want)


def test_c5_ultra_long_nanopore_ubam(sq):
    bam = synth.nanopore_ubam(5000, mean_length=20_000, max_length=1_000_000, seed=5)
    stream = bam[len(synth.bam_header()):]
    packed, recs, consumed, skipped = orc.decode_bam(stream)
    assert consumed == len(stream)
    oq, ons, oad = orc.QCMetrics(), orc.NanoStats(), orc.AdapterCounter(H.NANOPORE_ADAPTERS)
    oov = orc.OverrepresentedSequences()
    for o in (oq, ons, oad, oov):
        o.add(packed, recs)
    gq, gns, gad = sq.QCMetrics(), sq.NanoStats(), sq.AdapterCounter(H.NANOPORE_ADAPTERS)
    gov = sq.OverrepresentedSequences()
    n = 0
    for arr in sq.BamParser(io.BytesIO(bam), 8 << 20):
        for g in (gq, gns, gad, gov):
            g.add_record_array(arr)
        n += len(arr)
    assert n == len(recs) == 5000
    H.assert_same(H.dump_qc(gq), H.odump_qc(oq))
    H.assert_same(H.dump_nano(gns), H.odump_nano(ons))
    H.assert_same(H.dump_adapters(gad), H.odump_adapters(oad))
    H.assert_same(H.dump_overrep(gov), H.odump_overrep(oov))
