"""One rank of the multi-GPU parity check (launched by tests/test_multi_gpu.py through
torch.distributed.run, one process per GPU): every rank feeds its contiguous shard of the same
seeded FASTQ text to the CUDA collectors, the order-dependent tables are merged over NCCL
(sequali_b200.sharded), and the merged result must equal the CPU oracle's single pass over the
whole text, bit for bit."""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["SEQUALI_B200_DEVICE"] = str(local)
    import sequali_b200 as sq
    from sequali_b200 import sharded, synth
    from oracle import oracle as orc
    from tests import helpers as H
    comm = sharded.NcclComm.from_env()  # NCCL inside libsqgpu: no torch in this process
    sharded.use_comm(comm)
    assert "torch" not in sys.modules

    cases = [
        # (text, cut fractions, dedup kwargs, overrep kwargs, buffersize)
        (synth.illumina_fastq(40_000, length=150, seed=41, n_tiles=7),
         dict(max_stored_fingerprints=2000, front_sequence_offset=64, back_sequence_offset=0),
         dict(max_unique_fragments=3000, sample_every=3), 1 << 26),
        (synth.illumina_fastq(30_000, length=100, seed=42, n_tiles=5, variable_length=True),
         dict(max_stored_fingerprints=1000, front_sequence_offset=64, back_sequence_offset=0),
         dict(max_unique_fragments=10 ** 6, sample_every=8), 700_000),
        (synth.illumina_fastq(20_000, length=150, seed=43, n_tiles=6, tile_runs=False),
         dict(max_stored_fingerprints=100_000, front_sequence_offset=64, back_sequence_offset=0),
         dict(max_unique_fragments=60_000, sample_every=2), 1 << 26),
        # nanopore headers: NanoStats records of all ranks in read order, PerTileQuality switches itself off
        (synth.nanopore_fastq(400, mean_length=2500, max_length=30_000, seed=44),
         dict(max_stored_fingerprints=150, front_sequence_offset=64, back_sequence_offset=0),
         dict(max_unique_fragments=20_000, sample_every=1), 1 << 20),
    ]
    for ci, (text, dd_kw, ov_kw, bufsize) in enumerate(cases):
        adapters = H.NANOPORE_ADAPTERS if ci == 3 else H.ILLUMINA_ADAPTERS
        recs, _ = orc.parse_fastq(text)
        n = len(recs)
        cuts = [0] + [int(n * (g + 1) / world) + (7 if g + 1 < world else 0) for g in range(world)]
        lo, hi = cuts[rank], cuts[rank + 1]
        b0 = int(recs[lo]["name_off"]) - 1 if lo < n else len(text)
        b1 = int(recs[hi]["name_off"]) - 1 if hi < n else len(text)
        coll = sharded.ShardedCollectors(sq, adapters, first_record=lo, dedup_kwargs=dd_kw,
                                         overrep_kwargs=ov_kw)
        for arr in sq.FastqParser(io.BytesIO(text[b0:b1]), bufsize):
            coll.add_record_array(arr)
        got = coll.merge()
        if rank == world - 1:
            want = H.oracle_single_end(text, adapters, dedup_kwargs=dd_kw, overrep_kwargs=ov_kw)
            assert got["qc"]["base_count_table"].tolist() == want["qc"]["base"]
            assert got["qc"]["phred_count_table"].tolist() == want["qc"]["phred"]
            assert got["qc"]["gc_content"].tolist() == want["qc"]["gc"]
            assert got["qc"]["phred_scores"].tolist() == want["qc"]["mean_phred"]
            assert got["qc"]["end_anchored_base_count_table"].tolist() == want["qc"]["ea_base"]
            assert got["qc"]["end_anchored_phred_count_table"].tolist() == want["qc"]["ea_phred"]
            assert got["qc"]["number_of_reads"] == want["qc"]["number_of_reads"] == n
            assert got["qc"]["max_length"] == want["qc"]["max_length"]
            assert got["adapters_number_of_sequences"] == want["adapters"]["number_of_sequences"]
            nano = H.dump_nano(got["nano"])
            assert nano["infos"] == want["nano"]["infos"], f"case {ci}: NanoStats records"
            for k in ("number_of_reads", "minimum_time", "maximum_time"):
                assert nano[k] == want["nano"][k], (ci, k, nano[k], want["nano"][k])
            assert (nano["skipped_reason"] is None) == (want["nano"]["skipped_reason"] is None)
            for (a, f, r), (wa, wf, wr) in zip(got["adapters"], want["adapters"]["counts"]):
                assert a == wa and f.tolist() == wf and r.tolist() == wr
            want_tiles = want["ptq"]["tiles"]
            got_tiles = [(t, H.f64_bits(np.array(e, dtype=np.float64)), list(c)) for t, e, c in got["ptq"]["tiles"]]
            assert got_tiles == [(t, e, c) for t, e, c in want_tiles], f"case {ci}: per-tile sums differ"
            assert got["ptq"]["number_of_reads"] == want["ptq"]["number_of_reads"]
            assert sorted(got["dedup"]["counts"].tolist()) == want["dedup"]["sorted"], f"case {ci}: dedup"
            assert got["dedup"]["counts"].tolist() == want["dedup"]["slot_order"], f"case {ci}: dedup slot order"
            assert got["dedup"]["modulo_bits"] == want["dedup"]["modulo_bits"]
            assert got["dedup"]["tracked_sequences"] == want["dedup"]["tracked_sequences"]
            H.assert_same(H.dump_overrep(got["overrep"]), want["overrep"])
            print(f"case {ci}: merged tables of {world} ranks equal the oracle "
                  f"(dedup bits {want['dedup']['modulo_bits']}, unique fragments "
                  f"{want['overrep']['collected_unique_fragments']})", flush=True)
        comm.barrier()
    if rank == world - 1:
        print("MGPU PARITY OK", flush=True)
    comm.close()


if __name__ == "__main__":
    main()
