"""N>1 host logic on CPU: two gloo ranks each run the hot loop on a contiguous
shard (through the CPU oracle -- there is no GPU here) and merge their tables
with sequali_b200.sharded; the merged tables must equal the single-process
result on the whole input."""
import io
import os
import socket

import numpy as np
import pytest

from tests import helpers as H
from oracle import oracle as orc
from sequali_b200 import sharded, synth

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_tables(text, lo, hi, recs_all):
    recs = recs_all[lo:hi]
    buf = np.frombuffer(text, np.uint8)
    qc, ad, ptq, ns = orc.QCMetrics(), orc.AdapterCounter(H.ILLUMINA_ADAPTERS), orc.PerTileQuality(), orc.NanoStats()
    qc.add(buf, recs)
    ad.add(buf, recs)
    ptq.add(buf, recs)
    ns.add(buf, recs)
    return H.odump_qc(qc), H.odump_adapters(ad), H.odump_ptq(ptq), ns


def _worker(rank, world, port, text, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import sharded_adapters as A
    sharded.use_comm(A.TorchComm())
    try:
        recs, _ = orc.parse_fastq(text)
        # shard at a tile border so that PerTileQuality stays exact
        tiles = [bytes(text[int(r["name_off"]):int(r["name_off"]) + int(r["name_len"])]).split(b":")[4] for r in recs]
        cut = next(i for i in range(len(recs) // 2, len(recs)) if tiles[i] != tiles[i - 1])
        bounds = [(0, cut), (cut, len(recs))]
        lo, hi = bounds[rank]
        qc, ad, ptq, _ = _shard_tables(text, lo, hi, recs)
        m_qc = sharded.merge_qc(np.array(qc["base"], np.uint64), np.array(qc["phred"], np.uint64),
                                np.array(qc["ea_base"], np.uint64), np.array(qc["ea_phred"], np.uint64),
                                np.array(qc["gc"], np.uint64), np.array(qc["mean_phred"], np.uint64))
        m_ad = sharded.merge_adapter_counts([(a, np.array(f, np.uint64), np.array(r, np.uint64))
                                             for a, f, r in ad["counts"]])
        # error sums travel as bit patterns here; tiles do not straddle the cut, so nothing is added
        m_pt, straddling = sharded.merge_tile_counts(ptq["tiles"])
        assert sharded.allreduce_max(rank) == world - 1
        if rank == 0:
            out.put((m_qc, m_ad, m_pt, straddling))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            b = [sharded.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_two_rank_merge_equals_single_process():
    import torch.multiprocessing as mp
    text = synth.illumina_fastq(4000, length=100, seed=77, n_tiles=6, variable_length=True)
    recs, _ = orc.parse_fastq(text)
    want_qc, want_ad, want_pt, _ = _shard_tables(text, 0, len(recs), recs)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, text, out)) for r in range(2)]
    for p in procs:
        p.start()
    m_qc, m_ad, m_pt, straddling = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert m_qc["base_count_table"].tolist() == want_qc["base"]
    assert m_qc["phred_count_table"].tolist() == want_qc["phred"]
    assert m_qc["end_anchored_base_count_table"].tolist() == want_qc["ea_base"]
    assert m_qc["end_anchored_phred_count_table"].tolist() == want_qc["ea_phred"]
    assert m_qc["gc_content"].tolist() == want_qc["gc"]
    assert m_qc["phred_scores"].tolist() == want_qc["mean_phred"]
    for (a, f, r), (wa, wf, wr) in zip(m_ad, want_ad["counts"]):
        assert a == wa and f.tolist() == wf and r.tolist() == wr
    assert straddling == []
    assert [(t, e, c) for t, e, c in m_pt] == [(t, e, c) for t, e, c in want_pt["tiles"]]


# ------------------------------------------------------------------------------
# order-dependent collectors: the rank-to-rank protocol of sequali_b200.sharded
# (merge_dedup / merge_overrep / merge_pertile), driven through the oracle
# ------------------------------------------------------------------------------
def _exact_worker(rank, world, port, text, cuts, dd_kw, ov_kw, out):
    import torch.distributed as dist
    from tests import sharded_adapters as A
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded.use_comm(A.TorchComm())
    try:
        recs, _ = orc.parse_fastq(text)
        lo, hi = cuts[rank], cuts[rank + 1]
        dd = A.OracleDedup(deferred=rank > 0, **dd_kw)
        ov = A.OracleOverrep(deferred=rank > 0, first_record=lo, **ov_kw)
        pt = A.OraclePerTile()
        mid = (lo + hi) // 2
        for a, b in ((lo, mid), (mid, hi)):  # two record arrays per shard
            if b > a:
                dd.add(text, recs[a:b])
                ov.add(text, recs[a:b])
                pt.add(text, recs[a:b])
        counts, info = sharded.merge_dedup(dd)
        sharded.merge_overrep(ov)
        ptq = sharded.merge_pertile(pt, lo)
        got = dict(dd_counts=counts.tolist(), dd_info=info, ov=H.odump_overrep(ov.o), ptq=ptq)
        if rank == world - 1:  # every rank holds the merged result; take it from the last one
            out.put(got)
    finally:
        dist.destroy_process_group()


def _exact_case(text, world, cuts, dd_kw, ov_kw):
    import torch.multiprocessing as mp
    recs, _ = orc.parse_fastq(text)
    dd = orc.DedupEstimator(**dd_kw)
    ov = orc.OverrepresentedSequences(**ov_kw)
    pt = orc.PerTileQuality()
    dd.add(text, recs)
    ov.add(text, recs)
    rc = pt.add(text, recs)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exact_worker, args=(r, world, port, text, cuts, dd_kw, ov_kw, out))
             for r in range(world)]
    for p in procs:
        p.start()
    import queue
    got = None
    for _ in range(600):
        try:
            got = out.get(timeout=0.5)
            break
        except queue.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), "a rank died"
    assert got is not None
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["dd_info"]["modulo_bits"] == dd.info()["_modulo_bits"]
    assert got["dd_info"]["tracked_sequences"] == dd.info()["tracked_sequences"]
    assert got["dd_counts"] == dd.duplication_counts().tolist()        # slot order included
    H.assert_same(got["ov"], H.odump_overrep(ov))
    want_tiles = [(t, H.f64_bits(e), c.tolist()) for t, e, c in pt.get_tile_counts()]
    got_tiles = [(t, H.f64_bits(np.array(e)), list(c)) for t, e, c in got["ptq"]["tiles"]]
    assert got_tiles == want_tiles
    assert got["ptq"]["number_of_reads"] == pt.number_of_reads
    assert (got["ptq"]["skipped_record"] is not None) == (rc == 1)
    return dd, ov, got


@pytest.mark.parametrize("ov_kw", [dict(max_unique_fragments=400, sample_every=3),     # cap crossed in shard 0
                                   dict(max_unique_fragments=9000, sample_every=3),    # ... in shard 1
                                   dict(max_unique_fragments=10 ** 6, sample_every=8)])  # never
def test_two_rank_order_dependent_merges(ov_kw):
    # tiles in runs: the cut falls inside a run, so that tile straddles the border
    text = synth.illumina_fastq(3000, length=100, seed=21, n_tiles=5, variable_length=True)
    dd_kw = dict(max_stored_fingerprints=300, front_sequence_offset=64, back_sequence_offset=0)
    dd, ov, _ = _exact_case(text, 2, [0, 1333, 3000], dd_kw, ov_kw)
    assert dd.info()["_modulo_bits"] >= 2  # escalations happened in both shards


def test_three_rank_random_tiles_and_unparsable_header():
    # every tile on every rank; the module switches itself off inside the second shard
    text = synth.illumina_fastq(1500, length=60, seed=22, n_tiles=4, tile_runs=False)
    recs, _ = orc.parse_fastq(text)
    no = int(recs[900]["name_off"])
    text = text[:no] + text[no:].replace(b":", b"_", 5)  # record 900 loses its tile field
    dd_kw = dict(max_stored_fingerprints=200, front_sequence_offset=64, back_sequence_offset=0)
    _, _, got = _exact_case(text, 3, [0, 501, 1007, 1500], dd_kw,
                            dict(max_unique_fragments=1500, sample_every=2))
    assert got["ptq"]["skipped_record"] == 900 and got["ptq"]["number_of_reads"] == 900
