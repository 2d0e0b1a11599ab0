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
