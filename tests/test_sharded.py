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
