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


def paired_case(sq, sharded, synth, orc, H, rank, world):
    """Paired end over the ranks: InsertSizeMetrics (histogram + the two capped first-come adapter tables),
    the pair fingerprints of DedupEstimator, both read sides of QCMetrics / PerTileQuality / Overrepresented."""
    t1, t2 = synth.paired_fastq(24_000, seed=45)
    r1, _ = orc.parse_fastq(t1)
    r2, _ = orc.parse_fastq(t2)
    n = len(r1)
    cuts = [0] + [int(n * (g + 1) / world) + (5 if g + 1 < world else 0) for g in range(world)]
    lo, hi = cuts[rank], cuts[rank + 1]

    def piece(text, recs):
        b0 = int(recs[lo]["name_off"]) - 1 if lo < n else len(text)
        b1 = int(recs[hi]["name_off"]) - 1 if hi < n else len(text)
        return text[b0:b1]
    dd_kw = dict(max_stored_fingerprints=3000, front_sequence_offset=0, back_sequence_offset=0)
    for max_adapters in (10_000, 300):  # never full / full inside the first shard
        coll = sharded.ShardedPairedCollectors(sq, first_record=lo, dedup_kwargs=dd_kw)
        coll.ins = sharded.GpuInsert(sq.InsertSizeMetrics(max_adapters), rank > 0)
        rd1 = sq.FastqParser(io.BytesIO(piece(t1, r1)), 600_000)
        rd2 = sq.FastqParser(io.BytesIO(piece(t2, r2)), 600_000)
        for a in rd1:
            b = rd2.read(len(a))
            assert a.is_mate(b)
            coll.add_record_array_pair(a, b)
        got = coll.merge()
        if rank == world - 1:
            want = paired_oracle(orc, H, t1, t2, dd_kw, max_adapters)
            ins, wi = got["insert"], want["insert"]
            assert ins["sizes"].tolist() == wi["sizes"], "insert sizes"
            assert ins["adapters1"] == wi["adapters1_slot_order"] and ins["adapters2"] == wi["adapters2_slot_order"], \
                f"adapter tables (max_adapters={max_adapters})"
            for k in ("total_reads", "number_of_adapters_read1", "number_of_adapters_read2"):
                assert ins[k] == wi[k], (k, ins[k], wi[k])
            assert got["dedup"]["counts"].tolist() == want["dedup"]["slot_order"], "pair dedup"
            for side in ("1", "2"):
                assert got["qc" + side]["base_count_table"].tolist() == want["qc" + side]["base"]
                assert got["qc" + side]["phred_count_table"].tolist() == want["qc" + side]["phred"]
                assert got["qc" + side]["number_of_reads"] == n
                tiles = [(t, H.f64_bits(np.array(e, dtype=np.float64)), list(c)) for t, e, c in got["ptq" + side]["tiles"]]
                assert tiles == want["ptq" + side]["tiles"], "per-tile sums, read " + side
                H.assert_same(H.dump_overrep(got["overrep" + side]), want["overrep" + side])
            print(f"paired case (max_adapters={max_adapters}): merged tables of {world} ranks equal the oracle", flush=True)
        sharded.comm().barrier()


def paired_oracle(orc, H, t1, t2, dd_kw, max_adapters):
    r1, _ = orc.parse_fastq(t1)
    r2, _ = orc.parse_fastq(t2)
    b1, b2 = np.frombuffer(t1, np.uint8), np.frombuffer(t2, np.uint8)
    qc1, qc2, p1, p2 = orc.QCMetrics(), orc.QCMetrics(), orc.PerTileQuality(), orc.PerTileQuality()
    o1, o2 = orc.OverrepresentedSequences(), orc.OverrepresentedSequences()
    dd, ins = orc.DedupEstimator(**dd_kw), orc.InsertSizeMetrics(max_adapters)
    qc1.add(b1, r1); p1.add(b1, r1); o1.add(b1, r1)
    dd.add_pair(b1, r1, b2, r2)
    ins.add_pair(b1, r1, b2, r2)
    qc2.add(b2, r2); p2.add(b2, r2); o2.add(b2, r2)
    return dict(qc1=H.odump_qc(qc1), qc2=H.odump_qc(qc2), ptq1=H.odump_ptq(p1), ptq2=H.odump_ptq(p2),
                overrep1=H.odump_overrep(o1), overrep2=H.odump_overrep(o2), dedup=H.odump_dedup(dd),
                insert=H.odump_insert(ins))


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
    paired_case(sq, sharded, synth, orc, H, rank, world)
    comm.barrier()
    if rank == world - 1:
        print("MGPU PARITY OK", flush=True)
    comm.close()


if __name__ == "__main__":
    main()
