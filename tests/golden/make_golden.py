"""Generate the golden vectors under tests/golden/ with the UNMODIFIED reference
(oracle/_ref, built by oracle/build_ref.sh from /root/reference).

Run once in the build container:   python tests/golden/make_golden.py
The inputs are committed next to the outputs so that the fixtures do not depend
on the numpy version that produced them.  Doubles are stored as u64 bit
patterns.  The reference's own known-answer cases (tests/test_*.py of the
reference) that fit a table are restated in tests/test_oracle.py.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import helpers as H  # noqa: E402
from sequali_b200 import synth  # noqa: E402


def write(name, data: bytes):
    with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as f:
        f.write(data)


def dump(name, obj):
    blob = json.dumps(obj, separators=(",", ":"), sort_keys=True).encode()
    with gzip.GzipFile(os.path.join(HERE, name + ".gz"), "wb", mtime=0) as f:
        f.write(blob)


def main():
    ref = H.import_reference()
    assert ref is not None, "build oracle/_ref first (oracle/build_ref.sh)"

    # 94 error rates 10^-(q/10) as the reference's table holds them: one read of
    # one base per quality through PerTileQuality gives total_errors[0] = ERR[q]
    rates = []
    for q in range(94):
        p = ref.PerTileQuality()
        p.add_read(ref.FastqRecordView("a:b:c:d:1:x:y", "A", chr(33 + q)))
        rates.append(H.f64_bits(p.get_tile_counts()[0][1])[0])
    dump("error_rates.json", rates)

    illumina = synth.illumina_fastq(1500, length=100, seed=101, n_tiles=12)
    write("illumina_se.fastq.gz", illumina)
    dump("illumina_se.json", H.api_single_end(ref, illumina, H.ILLUMINA_ADAPTERS))

    # small tables force the order-dependent paths: dedup escalation and the fragment cap
    kw = dict(dedup_kwargs=dict(max_stored_fingerprints=120, front_sequence_offset=64,
                                back_sequence_offset=0),
              overrep_kwargs=dict(max_unique_fragments=400, sample_every=2))
    ragged = synth.illumina_fastq(1500, length=90, seed=102, n_tiles=30, tile_runs=False,
                                  variable_length=True)
    write("illumina_ragged.fastq.gz", ragged)
    dump("illumina_ragged.json", H.api_single_end(ref, ragged, H.ILLUMINA_ADAPTERS, **kw))

    r1, r2 = synth.paired_fastq(1000, length=100, seed=103)
    write("paired_R1.fastq.gz", r1)
    write("paired_R2.fastq.gz", r2)
    dump("paired.json", H.api_paired(ref, r1, r2))

    nano = synth.nanopore_fastq(40, mean_length=3000, max_length=15000, seed=104)
    write("nanopore.fastq.gz", nano)
    dump("nanopore.json", H.api_single_end(ref, nano, H.NANOPORE_ADAPTERS))

    bam = synth.nanopore_ubam(40, mean_length=2000, max_length=12000, seed=105)
    write("nanopore.bam.gz", bam)
    import io
    dump("nanopore_bam.json", H.api_single_end(ref, b"", H.NANOPORE_ADAPTERS,
                                               fileobj=io.BytesIO(bam), bam=True))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
