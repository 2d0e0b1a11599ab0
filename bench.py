#!/usr/bin/env python
"""bench.py -- all-module QC throughput of the sequali hot path on B200.

Metric (BASELINE.json): Gbases/s (and reads/s) of all default modules over
synthetic NovaSeq-style 150 bp single-end reads (recipe C2).

  python bench.py --gpus 1 --steps 5 --warmup 3          # this repo (CUDA path)
  python bench.py --impl reference --gpus 1 ...          # the reference's CPU path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot loop (src/sequali/__main__.py:279-306:
QCMetrics, PerTileQuality, OverrepresentedSequences, NanoStats, AdapterCounter,
DedupEstimator) over the full input with fresh collectors.

  value     input resident in HBM when the timed region starts (device parse +
            all collectors + table read-back), CUDA events on the launch stream
  e2e       the same loop with HOST text in pinned memory through the C ABI
            (sq_fastq_stream_*: windows copied H2D ahead of the parser on a
            copy stream) -> kernels -> getters (D2H); e2e.fileobj_api = through
            FastqParser(file object)
  roofline  dominant kernel: algorithmic bytes (record text, read once) per
            launch / its mean launch time (CUDA events), against MEASURED_PEAKS
  cpu_baseline  the unmodified reference (oracle/_ref) on one host core over a
            bounded prefix of the same input

  configs   the other workloads BASELINE.json names (C1 1 M reads, C3 paired end,
            C4 ultra-long nanopore FASTQ, C5 the same reads as unaligned BAM) at bounded
            sizes: Gbases/s through the reference-shaped API on host bytes, the unmodified
            reference on the identical bytes (1 thread), and the step-level roofline fraction

Multi-GPU: reads shard across ranks (each rank owns an equal, contiguous slice
of the record stream: weak scaling), no data-path collective; the tables are
merged exactly inside the timed region over NCCL (sq_comm_* of libsqgpu: no
torch in this process -- torchrun only launches the ranks).  Before timing, a
2 M-read sharded pass is compared with a single-rank pass over the same text
("parity_checked").
"""
from __future__ import annotations

import argparse
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ILLUMINA_ADAPTERS = ["AGATCGGAAGAG", "TGGAATTCTCGG", "GATCGTCGGACT", "CTGTCTCTTATA",
                     "GGGGGGGGGGGG", "AAAAAAAAAAAA"]
READ_LENGTH = 150
METRIC = "gbases_per_s_all_module_qc"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._halt = index, [], set(), threading.Event()
        self.max_mhz = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------
# the hot loop, shaped like src/sequali/__main__.py:214-306 (single end)
# ----------------------------------------------------------------------------
def make_modules(mod):
    return dict(
        qc=mod.QCMetrics(), ptq=mod.PerTileQuality(), ov=mod.OverrepresentedSequences(),
        ns=mod.NanoStats(), ad=mod.AdapterCounter(ILLUMINA_ADAPTERS),
        dd=mod.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0))


def feed(mods, arr):
    mods["qc"].add_record_array(arr)
    mods["ptq"].add_record_array(arr)
    mods["ov"].add_record_array(arr)
    mods["ns"].add_record_array(arr)
    mods["ad"].add_record_array(arr)
    mods["dd"].add_record_array(arr)


def read_results(mods):
    """The getters the report calls (report_modules.py:2537-2605); returns the
    additive tables as numpy arrays (merged across ranks) and result bytes."""
    qc = mods["qc"]
    t = [np.frombuffer(x, dtype=np.uint64) for x in (
        qc.base_count_table(), qc.phred_count_table(), qc.end_anchored_base_count_table(),
        qc.end_anchored_phred_count_table(), qc.gc_content(), qc.phred_scores())]
    for _, f, r in mods["ad"].get_counts():
        t += [np.frombuffer(f, dtype=np.uint64), np.frombuffer(r, dtype=np.uint64)]
    tiles = mods["ptq"].get_tile_counts()
    dups = mods["dd"].duplication_counts()
    ov = mods["ov"].overrepresented_sequences(threshold_fraction=0.001, min_threshold=100)
    nbytes = sum(x.nbytes for x in t) + len(dups) * 8 + len(tiles) * 16 * max(qc.max_length, 1)
    return np.concatenate(t), nbytes, dict(tiles=len(tiles), dups=len(dups), overrep=len(ov),
                                           reads=qc.number_of_reads)


class HostText(io.RawIOBase):
    """File-like over host memory: what a decompressor thread hands the parser."""

    def __init__(self, arr: np.ndarray):
        self._mv, self._pos = memoryview(arr), 0

    def readinto(self, b):
        n = min(len(b), len(self._mv) - self._pos)
        b[:n] = self._mv[self._pos:self._pos + n]
        self._pos += n
        return n

    def readable(self):
        return True


def ncu_traffic(kernel: str, chunk_reads: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed
    `ncu --set full` capture (profiles/*_traffic.json; captured on record arrays of 4 Mi reads, the
    default array size here).  (None, reason) when there is no capture for this kernel / array size."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files or chunk_reads != 1 << 22:
        return None, "no ncu capture for this record-array size"
    doc = json.load(open(files[-1]))
    alias = {"k_fused_columns_qhist": "k_fused_columns<", "k_fused_columns_bins": "k_fused_columns<",
             "k_fused_reads<NW>": "k_fused_reads<"}
    want = alias.get(kernel, kernel)
    for name, rec in doc["kernels"].items():
        if name.startswith(want):
            return int(rec["dram_bytes_per_launch"]), os.path.basename(files[-1]) + ": " + name
    return None, "kernel not in " + os.path.basename(files[-1])


NANOPORE_ADAPTERS = ["TTACGTATTGCT", "GCAATACGTAAC", "CTTGCGGGCGGC", "GGTAGTAGGTTC", "GAGGCGAGCGGT", "CAAGATACGCAC",
                     "GTGACTTGCCTG", "ATCGCCTACCGT", "TCTATCTTCTTT", "TCTTCAGAGGAG", "GATATTGCTGGG", "TGATATTGCTTT",
                     "GTACGTATTGCT", "ACGTAACTGAAC"]  # src/sequali/adapters/adapter_list.tsv:44-57


def loop_single_end(mod, fileobj, adapters, bam=False, buffersize=None):
    """src/sequali/__main__.py:214-306 (single end) + the getters the report calls; returns bases."""
    mods = dict(qc=mod.QCMetrics(), ptq=mod.PerTileQuality(), ov=mod.OverrepresentedSequences(), ns=mod.NanoStats(),
                ad=mod.AdapterCounter(adapters), dd=mod.DedupEstimator(front_sequence_offset=64, back_sequence_offset=0))
    kw = {} if buffersize is None else dict(initial_buffersize=buffersize)
    parser = (mod.BamParser if bam else mod.FastqParser)(fileobj, **kw)
    for arr in parser:
        feed(mods, arr)
    bases = int(np.frombuffer(mods["qc"].base_count_table(), dtype=np.uint64).sum())
    mods["qc"].phred_count_table()
    mods["ad"].get_counts()
    mods["ptq"].get_tile_counts()
    mods["dd"].duplication_counts()
    mods["ov"].overrepresented_sequences(threshold_fraction=0.001, min_threshold=100)
    sum(1 for _ in mods["ns"].nano_info_iterator())
    return bases


def loop_paired(mod, f1, f2, buffersize=None):
    """src/sequali/__main__.py:284-303 (paired end) + getters; returns bases."""
    kw = {} if buffersize is None else dict(initial_buffersize=buffersize)
    qc1, qc2, p1, p2 = mod.QCMetrics(), mod.QCMetrics(), mod.PerTileQuality(), mod.PerTileQuality()
    o1, o2 = mod.OverrepresentedSequences(), mod.OverrepresentedSequences()
    dd = mod.DedupEstimator(front_sequence_offset=0, back_sequence_offset=0)
    ins = mod.InsertSizeMetrics()
    rd1, rd2 = mod.FastqParser(f1, **kw), mod.FastqParser(f2, **kw)
    for a in rd1:
        qc1.add_record_array(a)
        p1.add_record_array(a)
        o1.add_record_array(a)
        b = rd2.read(len(a))
        if len(a) != len(b) or not a.is_mate(b):
            raise RuntimeError("mates out of step")
        dd.add_record_array_pair(a, b)
        ins.add_record_array_pair(a, b)
        qc2.add_record_array(b)
        p2.add_record_array(b)
        o2.add_record_array(b)
    bases = int(np.frombuffer(qc1.base_count_table(), dtype=np.uint64).sum() +
                np.frombuffer(qc2.base_count_table(), dtype=np.uint64).sum())
    for m in (p1, p2):
        m.get_tile_counts()
    for m in (o1, o2):
        m.overrepresented_sequences(threshold_fraction=0.001, min_threshold=100)
    dd.duplication_counts()
    ins.insert_sizes(), ins.adapters_read1(), ins.adapters_read2()
    return bases


def run_configs(gpu_mod, peak_gbs):
    """C1 / C3 / C4 / C5 of BASELINE.json at bounded sizes: the B200 build through the reference-shaped
    API on regular files in the page cache (file reads and device copies inside the timed region), the
    unmodified reference on the identical files with one thread, and text bytes / time / HBM peak for the
    B200 run."""
    from sequali_b200 import synth
    ref, _ = import_cpu_impl()
    big = 64 << 20

    def timed(fn, repeat):
        best, bases = None, 0
        for _ in range(repeat):
            t0 = time.perf_counter()
            bases = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, bases

    def entry(workload, nbytes, gpu_fn, cpu_fn):
        gpu_fn()  # warm-up (allocations, table growth)
        dt, bases = timed(gpu_fn, 2)
        row = {"workload": workload, "bases": bases, "text_bytes": int(nbytes),
               "value": round(bases / dt / 1e9, 3), "unit": "Gbases/s", "ms": round(dt * 1e3, 2),
               "path": "sequali._qc extension on a regular file in the page cache, open(path, 'rb') (read + H2D inside the "
                       "timed region)",
               "roofline": {"bound": "hbm", "frac_step": round(nbytes / dt / 1e9 / peak_gbs, 5),
                            "step_gbs": round(nbytes / dt / 1e9, 2), "peak": peak_gbs}}
        if ref is not None:
            cdt, cbases = timed(cpu_fn, 1)
            assert cbases == bases, (workload, cbases, bases)
            row["cpu_reference"] = {"value": round(cbases / cdt / 1e9, 4), "unit": "Gbases/s", "cores": 1,
                                    "kind": "reference", "sample": "the identical file, whole"}
            row["speedup_vs_1_core"] = round(cdt / dt, 1)
        return row

    import shutil
    import tempfile
    where = next((d for d in ("/dev/shm", tempfile.gettempdir())
                  if os.path.isdir(d) and shutil.disk_usage(d).free > (6 << 30)), None)
    tmp = tempfile.mkdtemp(prefix="sq_configs_", dir=where) if where else None

    def on_disk(name, data):
        """The input as a regular file (page cache): what the CLI hands the parsers for an uncompressed input.
        (No room for files on this box: the bytes themselves, read through an in-memory file object.)"""
        if tmp is None:
            return data
        path = os.path.join(tmp, name)
        with open(path, "wb") as f:
            f.write(data)
        return path

    def opened(src):
        return open(src, "rb") if isinstance(src, str) else io.BytesIO(src)

    def single(mod, path, adapters, bam=False, **kw):
        with opened(path) as f:
            return loop_single_end(mod, f, adapters, bam=bam, **kw)

    def paired(mod, p1, p2, **kw):
        with opened(p1) as f1, opened(p2) as f2:
            return loop_paired(mod, f1, f2, **kw)

    out = {}
    try:
        text = synth.illumina_fastq(1_000_000, READ_LENGTH, seed=1, n_tiles=192)
        path = on_disk("c1.fastq", text)
        out["C1"] = entry("1M synthetic Illumina 150 bp single-end reads, all default modules", len(text),
                          lambda: single(gpu_mod, path, ILLUMINA_ADAPTERS, buffersize=big),
                          lambda: single(ref, path, ILLUMINA_ADAPTERS))
        t1, t2 = synth.paired_fastq(400_000, seed=3)
        p1, p2 = on_disk("c3_1.fastq", t1), on_disk("c3_2.fastq", t2)
        out["C3"] = entry("paired-end 2x150 bp, 400k pairs (BASELINE: 50M pairs; scaled 1/125): adapter overlap "
                          "detection and InsertSizeMetrics", len(t1) + len(t2),
                          lambda: paired(gpu_mod, p1, p2, buffersize=big), lambda: paired(ref, p1, p2))
        del t1, t2
        text = synth.nanopore_fastq(20_000, mean_length=20_000, max_length=1_000_000, seed=4)
        p4 = on_disk("c4.fastq", text)
        out["C4"] = entry("synthetic ultra-long Nanopore reads (mean 20 kb, max 1 Mb), 20k reads (BASELINE: 10 Gbases; "
                          "scaled ~1/25) with guppy headers: NanoStats, 14 adapters, 21-mer overrepresentation", len(text),
                          lambda: single(gpu_mod, p4, NANOPORE_ADAPTERS, buffersize=big),
                          lambda: single(ref, p4, NANOPORE_ADAPTERS))
        bam = synth.nanopore_ubam(20_000, mean_length=20_000, max_length=1_000_000, seed=5)
        p5 = on_disk("c5.bam", bam)
        out["C5"] = entry("dorado-style unaligned BAM, 20k reads (BASELINE: 10 Gbases; scaled ~1/25) with channel / "
                          "duration tags via BamParser", len(bam),
                          lambda: single(gpu_mod, p5, NANOPORE_ADAPTERS, bam=True, buffersize=big),
                          lambda: single(ref, p5, NANOPORE_ADAPTERS, bam=True))
    finally:
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_next_rows():
    """SURVEY.md 8(f) rows beside the hot path, each against the unmodified reference on one core:
    seqident (8(f)4): Smith-Waterman identity of 21-mers against contaminant-sized targets, one launch for all
    pairs / one call per pair; bam_chain (8(f)3): the block_size chain of a short-read uBAM, host bytes in."""
    import ctypes as C
    import importlib
    from sequali_b200 import _lib, synth
    from sequali_b200._qc import _PinnedBuffer
    from sequali_b200.ext import _seqident
    out = {}
    rng = np.random.default_rng(11)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    pairs = []
    for k in range(20_000):
        target = letters[rng.integers(0, 4, size=int(rng.integers(200, 2000)))]
        at = int(rng.integers(0, len(target) - 21))
        q = target[at:at + 21].copy()
        q[int(rng.integers(0, 21))] = letters[int(rng.integers(0, 4))]
        pairs.append((target.tobytes().decode(), q.tobytes().decode()))
    cells = sum(len(t) * len(q) for t, q in pairs)
    _seqident.sequence_identities(pairs)  # warm-up: context of the extension module, its device buffers
    t0 = time.perf_counter()
    got = _seqident.sequence_identities(pairs)
    dt = time.perf_counter() - t0
    row = {"workload": "20k (target 200-2000 nt, 21-mer query) pairs, one launch, host strings in, floats out",
           "value": round(cells / dt / 1e9, 3), "unit": "Gcells/s", "ms": round(dt * 1e3, 2)}
    ref, _ = import_cpu_impl()
    if ref is not None:
        ref_si = importlib.import_module("sequali._seqident")
        t0 = time.perf_counter()
        want = [ref_si.sequence_identity(t, q) for t, q in pairs]
        cdt = time.perf_counter() - t0
        assert got == want
        row["cpu_reference"] = {"value": round(cells / cdt / 1e9, 3), "unit": "Gcells/s", "cores": 1, "kind": "reference",
                                "sample": "the identical pairs, one call per pair"}
        row["speedup_vs_1_core"] = round(cdt / dt, 1)
    out["seqident"] = row

    ctx = _lib.Context.get()
    n_rec, length, name_w, tags = 2_000_000, 150, 12, b"RGZrg1\0"
    body = 32 + name_w + (length + 1) // 2 + length + len(tags)
    rec = np.zeros((n_rec, 4 + body), dtype=np.uint8)
    import struct
    rec[:, :36] = np.frombuffer(struct.pack("<IiiBBHHHIiii", body, -1, -1, name_w, 0, 4680, 0, 4, length, -1, -1, 0), dtype=np.uint8)
    rec[:, 36:47] = np.frombuffer(np.char.zfill(np.arange(n_rec).astype("S11"), 11).tobytes(), dtype=np.uint8).reshape(n_rec, 11)
    rec[:, 48:48 + (length + 1) // 2] = rng.integers(0, 256, size=(n_rec, (length + 1) // 2), dtype=np.uint8)
    rec[:, 48 + (length + 1) // 2:48 + (length + 1) // 2 + length] = rng.integers(2, 41, size=(n_rec, length), dtype=np.uint8)
    rec[:, -len(tags):] = np.frombuffer(tags, dtype=np.uint8)
    n = rec.size
    pinned = _PinnedBuffer(ctx, n)
    C.memmove(pinned.ptr, rec.ctypes.data, n)
    del rec
    offs = np.zeros(n // 36 + 2, dtype=np.uint64)
    kept, skipped, used = C.c_uint64(), C.c_uint64(), C.c_uint64()
    res = {}
    for what in ("host", "device"):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            if what == "host":
                rc = ctx.lib.sq_bam_walk(pinned.ptr, n, offs.ctypes.data, len(offs), C.byref(kept), C.byref(skipped), C.byref(used))
            else:
                rc = ctx.lib.sq_bam_walk_device(ctx.h, pinned.ptr, n, 0, offs.ctypes.data, len(offs), C.byref(kept),
                                                C.byref(skipped), C.byref(used))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert rc == 0
        res[what] = (best, kept.value, used.value, int(offs[:kept.value].sum()))
    assert res["host"][1:] == res["device"][1:] and res["host"][1] == n_rec
    out["bam_chain"] = {"workload": f"block_size chain of {n_rec} unaligned 150 bp BAM records ({n >> 20} MiB, pinned host bytes)",
                        "value": round(n / res["device"][0] / 1e9, 2), "unit": "GB/s",
                        "path": "sq_bam_walk_device: H2D copy + candidate test + pointer doubling + offsets back to the host",
                        "cpu_walk": {"value": round(n / res["host"][0] / 1e9, 2), "unit": "GB/s", "cores": 1,
                                     "kind": "port", "sample": "the identical bytes (sq_bam_walk: the reference's loop, :1623-1637)"}}
    out["report"] = report_row()
    return out


_REF_REPORT = r'''
import json, sys, time
from sequali import report_modules as rm
from sequali._qc import FastqParser, QCMetrics, NanoStats
metrics, nano = QCMetrics(), NanoStats()
with open(sys.argv[1], "rb") as f:
    for arr in FastqParser(f):
        metrics.add_record_array(arr)
        nano.add_record_array(arr)
ml = metrics.max_length
ranges = list(rm.logarithmic_ranges(ml)) if ml > 500 else list(rm.equidistant_ranges(ml, 200))
t0 = time.perf_counter()
mods = rm.qc_metrics_modules(metrics, ranges)
t1 = time.perf_counter()
ns = rm.NanoStatsReport.from_nanostats(nano)
t2 = time.perf_counter()
print(json.dumps({"qc_ms": (t1 - t0) * 1e3, "nano_ms": (t2 - t1) * 1e3, "n50": mods[1].n50, "total_bases": mods[0].total_bases,
                  "time_reads": sum(ns.time_reads), "channels": len(ns.per_channel_bases)}))
'''


def report_row():
    """8(f)2: the aggregation the report does on the collectors' results -- qc_metrics_modules' table sums and length
    walk (report_modules.py:2537-2605) and NanoStatsReport.from_nanostats' loop over every read (:1952-2046) --
    on the device tables (sequali_b200.report) against the reference's Python on one core, same input."""
    import shutil
    import tempfile
    import sequali_b200.ext as sqx
    from sequali_b200 import report, synth
    text = synth.nanopore_fastq(100_000, mean_length=300, max_length=1_000_000, seed=12)
    big = synth.nanopore_fastq(1, mean_length=900_000, max_length=1_000_000, seed=13)  # one read near 1 Mb: long tables
    text = text + big
    metrics, nano = sqx.QCMetrics(), sqx.NanoStats()
    for arr in sqx.FastqParser(io.BytesIO(text), 64 << 20):
        metrics.add_record_array(arr)
        nano.add_record_array(arr)
    ranges = report.data_ranges_for(metrics.max_length)
    report.qc_metrics_tables(metrics, ranges), report.nanostats_report(nano)  # warm-up
    t0 = time.perf_counter()
    qc = report.qc_metrics_tables(metrics, ranges)
    t1 = time.perf_counter()
    ns = report.nanostats_report(nano)
    t2 = time.perf_counter()
    row = {"workload": f"{nano.number_of_reads} nanopore reads, longest {metrics.max_length} nt ({len(ranges)} position ranges): "
                       "table sums + length walk of qc_metrics_modules, per-read loop of NanoStatsReport.from_nanostats",
           "qc_tables_ms": round((t1 - t0) * 1e3, 2), "nanostats_ms": round((t2 - t1) * 1e3, 2),
           "path": "sequali_b200.report on the collectors' device tables (results to Python objects included)"}
    pkg_src = os.path.join(ROOT, "oracle", "_ref", "pkg_src")
    ref_dir = os.path.join(ROOT, "oracle", "_ref", "sequali")
    if os.path.exists(os.path.join(pkg_src, "report_modules.py")) and os.path.exists(os.path.join(ref_dir, "_qc.abi3.so")):
        tmp = tempfile.mkdtemp(prefix="sq_report_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            shutil.copytree(pkg_src, os.path.join(tmp, "sequali"))
            for so in os.listdir(ref_dir):
                if so.endswith(".so"):
                    shutil.copy(os.path.join(ref_dir, so), os.path.join(tmp, "sequali", so))
            with open(os.path.join(tmp, "in.fastq"), "wb") as f:
                f.write(text)
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([tmp, os.path.join(ROOT, "oracle", "_ref", "tests", "shim")])
            proc = subprocess.run([sys.executable, "-c", _REF_REPORT, os.path.join(tmp, "in.fastq")], env=env,
                                  capture_output=True, text=True, timeout=900)
            if proc.returncode == 0:
                ref = json.loads(proc.stdout)
                assert ref["n50"] == qc["sequence_length_distribution"]["n50"] and ref["total_bases"] == qc["summary"]["total_bases"]
                assert ref["time_reads"] == sum(ns["time_reads"]) and ref["channels"] == len(ns["per_channel_bases"])
                row["cpu_reference"] = {"qc_tables_ms": round(ref["qc_ms"], 1), "nanostats_ms": round(ref["nano_ms"], 1), "cores": 1,
                                        "kind": "reference", "sample": "the identical reads; report_modules.py unchanged "
                                        "(pygal stubbed: nothing is plotted in either arm)"}
                row["speedup_vs_1_core"] = {"qc_tables": round(ref["qc_ms"] / max(row["qc_tables_ms"], 1e-3), 1),
                                            "nanostats": round(ref["nano_ms"] / max(row["nanostats_ms"], 1e-3), 1)}
            else:
                row["cpu_reference"] = {"unavailable": proc.stderr.strip().splitlines()[-1][:200] if proc.stderr.strip() else "failed"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return row


def _bgzf_worker(chunk):
    from sequali_b200 import synth
    return synth.bgzf_compress(bytes(chunk), level=6, eof_marker=False)


def e2e_bgzf(sq, hostq, HostFastq, feed, make_modules, read_results, ctx, args):
    """End to end from BGZF-compressed HOST bytes: a prefix of the e2e text is compressed here with zlib
    level 6 in 65 280-byte members (bgzip's layout; one process per core, outside the timed region), then
    the timed loop is the reader of sq_fastq_stream_create_bgzf + the collectors + the getters."""
    import multiprocessing as mp
    text = np.frombuffer(hostq.view(), dtype=np.uint8)
    limit = min(len(text), args.bgzf_bytes)
    limit = int(np.flatnonzero(text[:limit] == 10)[-1]) + 1 if limit else 0      # whole lines ...
    text = text[:limit]
    n_lines = int((text == 10).sum())
    reads = n_lines // 4
    if n_lines % 4:                                                              # ... and whole records
        ends = np.flatnonzero(text == 10)
        text = text[:int(ends[reads * 4 - 1]) + 1]
    step = 65280 * 256
    chunks = [text[i:i + step].tobytes() for i in range(0, len(text), step)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 32)) as pool:
        parts = pool.map(_bgzf_worker, chunks)
    comp = b"".join(parts)
    compress_s = time.perf_counter() - t0
    hz = HostFastq.from_bytes(comp)

    def step_fn():
        mods = make_modules(sq)
        for arr in hz.record_arrays_bgzf(args.e2e_window):
            feed(mods, arr)
        _, nbytes, summary = read_results(mods)
        assert summary["reads"] == reads, (summary["reads"], reads)
        return nbytes

    step_fn()
    ctx.sync()
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        out_bytes = step_fn()
        ctx.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ctx.profile(True)
    step_fn()
    prof = ctx.profile_report()
    ctx.profile(False)
    inflate_ms = prof.get("k_bgzf_inflate", (0, 0.0))[1]
    hz.free()
    return {"value": round(reads * READ_LENGTH / best / 1e9, 4), "unit": "Gbases/s", "reads": reads,
            "inflate_kernel": {"ms": round(inflate_ms, 3), "text_gbs": round(len(text) / max(inflate_ms, 1e-9) / 1e6, 1),
                               "launches": prof.get("k_bgzf_inflate", (0, 0.0))[0]},
            "text_bytes": int(len(text)), "h2d_bytes_per_step": int(len(comp)), "d2h_bytes_per_step": int(out_bytes),
            "compression_ratio": round(len(text) / len(comp), 3), "text_gbs": round(len(text) / best / 1e9, 2),
            "h2d_gbs": round(len(comp) / best / 1e9, 2), "host_compress_s_outside_timed_region": round(compress_s, 2),
            "path": "pinned host BGZF bytes (zlib level 6, 65280-byte members) -> H2D of the compressed windows on a copy "
                    "stream -> k_bgzf_inflate (one warp per member) straight into the record array -> parser -> "
                    "sq_fused_add -> getters"}


def sharded_parity_check(sq, sharded, DeviceFastq, comm, rank, world, total_reads):
    """Every rank runs its shard of a `total_reads` stream through ShardedCollectors + merge(), then the
    whole stream alone with plain collectors; every table must be equal (doubles by bit pattern, the
    dedup counts in slot order).  Returns True, or raises with the first table that differs."""
    per = max(total_reads // world, 1)
    total = per * world
    shard = DeviceFastq.synth_illumina(per, READ_LENGTH, seed=2, chunk_reads=1 << 20, first_read=rank * per,
                                       total_reads=total)
    coll = sharded.ShardedCollectors(sq, ILLUMINA_ADAPTERS, first_record=rank * per)
    for arr in shard.record_arrays():
        coll.add_record_array(arr)
    got = coll.merge()
    got_ov = got["overrep"]
    merged = {
        **{"qc." + k: np.asarray(v).tolist() for k, v in got["qc"].items()},
        "adapters": [(a, f.tolist(), r.tolist()) for a, f, r in got["adapters"]],
        "ptq.tiles": [(int(t), np.asarray(e, dtype=np.float64).view(np.uint64).tolist(), [int(x) for x in c])
                      for t, e, c in got["ptq"]["tiles"]],
        "ptq.number_of_reads": int(got["ptq"]["number_of_reads"]),
        "dedup.counts": np.asarray(got["dedup"]["counts"]).tolist(),
        "dedup.info": (int(got["dedup"]["modulo_bits"]), int(got["dedup"]["tracked_sequences"])),
        "overrep.counts": dict(got_ov.sequence_counts()),
        "overrep.counters": (got_ov.number_of_sequences, got_ov.sampled_sequences, got_ov.collected_unique_fragments,
                             got_ov.total_fragments),
    }
    del coll, got, got_ov
    shard.free()
    whole = DeviceFastq.synth_illumina(total, READ_LENGTH, seed=2, chunk_reads=1 << 20, first_read=0, total_reads=total)
    sharded.use_comm(None)  # a plain single-rank pass
    try:
        mods = make_modules(sq)
        for arr in whole.record_arrays():
            feed(mods, arr)
        qc = mods["qc"]
        single = {
            **{"qc." + k: list(getattr(qc, k)()) for k in (
                "base_count_table", "phred_count_table", "end_anchored_base_count_table",
                "end_anchored_phred_count_table", "gc_content", "phred_scores")},
            "qc.number_of_reads": qc.number_of_reads, "qc.max_length": qc.max_length,
            "adapters": [(a, list(f), list(r)) for a, f, r in mods["ad"].get_counts()],
            "ptq.tiles": [(int(t), np.asarray(e, dtype=np.float64).view(np.uint64).tolist(), [int(x) for x in c])
                          for t, e, c in mods["ptq"].get_tile_counts()],
            "ptq.number_of_reads": mods["ptq"].number_of_reads,
            "dedup.counts": list(mods["dd"].duplication_counts()),
            "dedup.info": (mods["dd"]._modulo_bits, mods["dd"].tracked_sequences),
            "overrep.counts": dict(mods["ov"].sequence_counts()),
            "overrep.counters": (mods["ov"].number_of_sequences, mods["ov"].sampled_sequences,
                                 mods["ov"].collected_unique_fragments, mods["ov"].total_fragments),
        }
        del mods
    finally:
        sharded.use_comm(comm)
        whole.free()
    bad = [k for k in single if single[k] != merged.get(k)]
    n_bad = int(comm.allreduce_host_u64([len(bad)], "sum")[0])
    if n_bad:
        raise AssertionError(f"rank {rank}: sharded pass differs from the single-rank pass in {bad}")
    return True


# ----------------------------------------------------------------------------
def run_cuda(args):
    import sequali_b200 as sq
    from sequali_b200 import _lib
    from sequali_b200.device import DeviceFastq

    from sequali_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("SEQUALI_B200_DEVICE", str(local))
    ctx = _lib.Context.get()
    comm = sharded.NcclComm.from_env(ctx)  # NCCL inside libsqgpu; None for a single rank
    sharded.use_comm(comm)

    def max_over_ranks(x: float) -> float:
        if comm is None:
            return x
        bits = np.array([x], dtype=np.float64).view(np.uint64)  # positive doubles order like their bit patterns
        return float(comm.allreduce_host_u64(bits, "max").view(np.float64)[0])

    n_reads = args.reads  # per GPU (weak scaling)
    data = DeviceFastq.synth_illumina(n_reads, READ_LENGTH, seed=2, chunk_reads=args.chunk_reads,
                                      first_read=rank * n_reads, total_reads=n_reads * world)
    text_bytes = data.nbytes
    bases = n_reads * READ_LENGTH

    def barrier():
        ctx.sync()
        if comm is not None:
            comm.barrier()

    def merge(tables: np.ndarray):
        return tables  # single rank: nothing to merge (N > 1 goes through step_sharded)

    sharded_ms = {}  # rank 0's host view of the last sharded step

    def step_sharded(record_arrays, first_record):
        """N > 1: every rank runs the hot loop on its contiguous shard; the merge then makes every
        table what one sequential pass over all shards gives (sequali_b200.sharded: all-reduce of
        the additive tables, border-tile records forwarded to the tile's owner, dedup hashes handed
        to the table's owner, the overrepresented table travelling until full, then frozen-key counts)."""
        t_begin = time.perf_counter()
        coll = sharded.ShardedCollectors(sq, ILLUMINA_ADAPTERS, first_record=first_record)
        for arr in record_arrays:
            coll.add_record_array(arr)
        t_feed = time.perf_counter()
        out = coll.merge()
        ov = out["overrep"].overrepresented_sequences(threshold_fraction=0.001, min_threshold=100)
        sharded_ms.clear()
        sharded_ms.update({"feed (host, kernels still running)": round((t_feed - t_begin) * 1e3, 2), **coll.merge_ms,
                           "total": round((time.perf_counter() - t_begin) * 1e3, 2)})
        width = out["ptq"]["max_length"]
        nbytes = (sum(v.nbytes for v in out["qc"].values() if isinstance(v, np.ndarray)) +
                  sum(f.nbytes + r.nbytes for _, f, r in out["adapters"]) +
                  len(out["dedup"]["counts"]) * 8 + len(out["ptq"]["tiles"]) * 16 * max(width, 1))
        return nbytes, dict(tiles=len(out["ptq"]["tiles"]), dups=len(out["dedup"]["counts"]), overrep=len(ov),
                            reads=out["qc"]["number_of_reads"])

    def step_resident():
        if world > 1:
            return step_sharded(data.record_arrays(), rank * n_reads)
        mods = make_modules(sq)
        for arr in data.record_arrays():
            feed(mods, arr)
        tables, nbytes, summary = read_results(mods)
        merge(tables)
        return nbytes, summary

    # ---- N > 1: the sharded pass equals a single-rank pass over the same text ---------------
    parity_checked = None
    if world > 1:
        parity_checked = sharded_parity_check(sq, sharded, DeviceFastq, comm, rank, world, args.parity_reads)

    # ---- HBM-resident timing -------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ctx.timer_start()
    t0 = time.perf_counter()
    step_ends = []
    for _ in range(args.steps):
        d2h, summary = step_resident()
        step_ends.append(time.perf_counter())  # (a step ends with its getters: the host has waited for the device)
    barrier()
    ms = ctx.timer_stop()
    ms_each_step = [round((b - a) * 1e3, 2) for a, b in zip([t0] + step_ends[:-1], step_ends)]
    if world > 1:
        print(f"[bench] rank {rank}: ms of each timed step (host clock): {ms_each_step}", file=sys.stderr, flush=True)
    wall = time.perf_counter() - t0
    sharded_resident = dict(sharded_ms)
    clocks = sampler.finish()
    launches = ctx.launch_count - launches0
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * bases / (ms_per_step * 1e-3) / 1e9

    # ---- where one step's wall time goes (host view; outside the timed region) ---
    t0 = time.perf_counter()
    mods = make_modules(sq)
    t1 = time.perf_counter()
    arrays = list(data.record_arrays())
    ctx.sync()
    t2 = time.perf_counter()
    per_mod = {}
    for key in ("qc", "ptq", "ov", "ns", "ad", "dd"):
        ta = time.perf_counter()
        for arr in arrays:
            mods[key].add_record_array(arr)
        ctx.sync()
        per_mod[key] = round((time.perf_counter() - ta) * 1e3, 3)
    t3 = time.perf_counter()
    getter_ms = {}
    for key, fn in (("qc_tables", lambda: [mods["qc"].base_count_table(), mods["qc"].phred_count_table(),
                                           mods["qc"].end_anchored_base_count_table(),
                                           mods["qc"].end_anchored_phred_count_table(),
                                           mods["qc"].gc_content(), mods["qc"].phred_scores()]),
                    ("adapter_counts", lambda: mods["ad"].get_counts()),
                    ("tile_counts", lambda: mods["ptq"].get_tile_counts()),
                    ("duplication_counts", lambda: mods["dd"].duplication_counts()),
                    ("overrepresented", lambda: mods["ov"].overrepresented_sequences(
                        threshold_fraction=0.001, min_threshold=100))):
        ta = time.perf_counter()
        fn()
        getter_ms[key] = round((time.perf_counter() - ta) * 1e3, 3)
    t4 = time.perf_counter()
    del arrays, mods
    host_ms = {"create_modules": round((t1 - t0) * 1e3, 3), "parse": round((t2 - t1) * 1e3, 3),
               "add_record_array": per_mod, "getters": round((t4 - t3) * 1e3, 3),
               "getters_each": getter_ms}

    # ---- per-kernel times of one step (CUDA events around every launch) --------
    # (profiled pass: everything on the launch stream, one host thread -- with the read-ahead thread enqueueing the
    # next array's scan on the same stream the gaps between kernels would measure the interleaving, not the idle time)
    os.environ["SEQUALI_B200_NO_PREFETCH"] = "1"
    ctx.profile(True)
    step_resident()
    prof = ctx.profile_report()
    ctx.profile(False)
    del os.environ["SEQUALI_B200_NO_PREFETCH"]
    gaps = {k[4:]: v for k, v in prof.items() if k.startswith("gap>")}
    prof = {k: v for k, v in prof.items() if not k.startswith("gap>")}
    kernel_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    peak, peak_src = measured_peak_gbs()
    n_chunks = len(data.chunks)
    top_launch_bytes = text_bytes / n_chunks  # one launch per chunk reads that chunk's text once
    top_ms_per_launch = top[1][1] / top[1][0]
    achieved = top_launch_bytes / (top_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(top[0], args.chunk_reads)
    step_gbs = text_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {
        # the number that answers north_star: text bytes of the whole step (read once) / step time / peak
        "frac_step": round(step_gbs / peak, 4), "step_gbs": round(step_gbs, 1),
        "bound": "hbm", "kernel": top[0], "achieved": round(achieved, 1), "peak": peak,
        "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(top_launch_bytes),
        "launches_per_step": top[1][0], "ms_per_launch": round(top_ms_per_launch, 4),
        "share_of_kernel_time": round(top[1][1] / kernel_ms, 3),
        "all_kernels_gbs": round(text_bytes / (kernel_ms * 1e-3) / 1e9, 1),
        "all_kernels_frac": round(text_bytes / (kernel_ms * 1e-3) / 1e9 / peak, 4),
        "kernels_ms_per_step": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:48]},
        # stream time between launches (memsets, copies, host syncs), by the kernel that follows
        "gaps_ms_per_step": {k: round(v[1], 3) for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:16]},
        "gaps_ms_total": round(sum(v[1] for v in gaps.values()), 3),
    }

    # ---- end to end: host text -> parser -> collectors -> getters --------------
    e2e = None
    if not args.no_e2e:
        from sequali_b200.device import HostFastq
        e2e_reads = min(n_reads, args.e2e_reads)
        hostq, e2e_reads = HostFastq.from_device(data, e2e_reads)

        def timed(step_fn, steps):
            step_fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                out = step_fn()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / steps), out

        # (1) the C-ABI with HOST buffers: pinned host text -> sq_fastq_stream_next (windows copied H2D on a
        #     copy stream ahead of the parser, inside the timed region) -> sq_fused_add -> getters (D2H)
        def step_e2e():
            if world > 1:
                return step_sharded(hostq.record_arrays(args.e2e_window), rank * e2e_reads)[0]
            mods = make_modules(sq)
            for arr in hostq.record_arrays(args.e2e_window):
                feed(mods, arr)
            tables, nbytes, _ = read_results(mods)
            merge(tables)
            return nbytes

        e2e_steps = max(1, min(args.steps, 3))
        dt, out_bytes = timed(step_e2e, e2e_steps)
        e2e = {"value": round(world * e2e_reads * READ_LENGTH / dt / 1e9, 4), "unit": "Gbases/s",
               "h2d_bytes_per_step": int(hostq.nbytes), "d2h_bytes_per_step": int(out_bytes),
               "reads_per_step_per_gpu": int(e2e_reads), "steps": e2e_steps, "window_bytes": args.e2e_window,
               "h2d_gbs": round(hostq.nbytes / dt / 1e9, 2),
               "path": "pinned host text -> sq_fastq_stream_next (triple-buffered cudaMemcpyAsync H2D on a copy "
                       "stream, overlapped with the kernels) -> sq_fused_add -> getters"}

        # (2) the reference-shaped Python API with a host FILE OBJECT (readinto into pinned staging first:
        #     one extra host copy by the Python file object, as with the reference's xopen stream)
        host = np.frombuffer(hostq.view(), dtype=np.uint8)

        import sequali_b200.ext as sqx  # the CPython extension `_qc`: what `import sequali` gives a user

        def step_fileobj():
            mods = make_modules(sqx)
            for arr in sqx.FastqParser(HostText(host), args.buffersize):
                feed(mods, arr)
            tables, nbytes, _ = read_results(mods)
            return nbytes

        if world == 1:  # (single rank only: the sharded loop is measured by `value` and `e2e`)
            dt2, _ = timed(step_fileobj, 1)
            e2e["fileobj_api"] = {"value": round(e2e_reads * READ_LENGTH / dt2 / 1e9, 4), "unit": "Gbases/s",
                                  "buffersize": args.buffersize,
                                  "path": "sequali._qc extension: FastqParser(in-memory file object).readinto (one host "
                                          "memcpy per byte, single thread) -> pinned staging -> H2D -> kernels -> getters"}
        # (2b) the same through a REGULAR FILE, open(path, "rb"), what the CLI hands the parser for an uncompressed
        #      input: the extension fetches the bytes with pread() from several threads (page cache -> pinned staging)
        if world == 1:
            import shutil
            import tempfile
            where = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            room = shutil.disk_usage(where).free
            n_file = len(host) if room > len(host) + (2 << 30) else 0
            if n_file:
                fd, path = tempfile.mkstemp(prefix="sq_bench_", suffix=".fastq", dir=where)
                try:
                    with os.fdopen(fd, "wb") as f:
                        f.write(memoryview(host))

                    def step_file():
                        mods = make_modules(sqx)
                        with open(path, "rb") as f:
                            for arr in sqx.FastqParser(f, args.buffersize):
                                feed(mods, arr)
                        return read_results(mods)[1]

                    dt3, _ = timed(step_file, 2)
                    e2e["fileobj_api"]["regular_file"] = {
                        "value": round(e2e_reads * READ_LENGTH / dt3 / 1e9, 4), "unit": "Gbases/s",
                        "file_gbs": round(n_file / dt3 / 1e9, 2),
                        "path": f"open(path, 'rb') on {where} (page cache) -> FastqParser: pread from up to 8 threads into "
                                "pinned staging, one record array read ahead -> H2D -> kernels -> getters"}
                finally:
                    os.unlink(path)
        del host

        # (3) the same host text as BGZF members (what bgzip writes): the compressed bytes cross PCIe and are
        #     inflated on the device (sq_fastq_stream_create_bgzf, csrc/inflate.cu) -- SURVEY.md 8(f)1
        if world == 1:
            e2e["bgzf"] = e2e_bgzf(sq, hostq, HostFastq, feed, make_modules, read_results, ctx, args)
        hostq.free()

    # ---- CPU baseline: the unmodified reference on one core, bounded sample ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(data, args.cpu_reads)

    # ---- the other configurations BASELINE.json names (bounded sizes, single GPU) -------------
    configs = next_rows = None
    if rank == 0 and world == 1 and not args.no_configs:
        import sequali_b200.ext as sqx
        data.free()  # the 100 M reads leave HBM first
        configs = run_configs(sqx, measured_peak_gbs()[0])
        next_rows = run_next_rows()

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": "Gbases/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
            "ms_each_step_rank0_host_clock": ms_each_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 (+f64 ordered sums)",
            "data": "synthetic",
            "config": {"workload": f"{n_reads} synthetic NovaSeq {READ_LENGTH} bp single-end reads per GPU, "
                                   "936 tiles in runs, all default modules (QCMetrics, PerTileQuality, "
                                   "OverrepresentedSequences, NanoStats, AdapterCounter, DedupEstimator)",
                       "reads_per_gpu": n_reads, "read_length": READ_LENGTH,
                       "text_bytes_per_gpu": int(text_bytes), "record_arrays_per_step": n_chunks,
                       "numa_node_of_rank0": ctx.numa_node,
                       "l2": "inputs larger than L2" if text_bytes > 200e6 else "input smaller than L2",
                       "parallelism": f"{world} x contiguous read shards" + (
                           "" if world == 1 else ", exact merges over NCCL inside libsqgpu (all-reduce of the additive "
                           "tables on the device; border-tile records, dedup hashes and the overrepresented table "
                           "exchanged between ranks)")},
            "reads_per_s": round(world * n_reads / (ms_per_step * 1e-3), 1),
            "wall_ms_per_step": round(wall * 1e3 / args.steps, 3),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            **({"parity_checked": parity_checked} if parity_checked is not None else {}),
            **({"sharded_ms_last_step_rank0": sharded_resident} if world > 1 else {}),
            "result_summary": summary, "host_ms_one_step": host_ms,
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if configs:
            line["configs"] = configs
        if next_rows:
            line["next_rows"] = next_rows
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.barrier()
        comm.close()


# ----------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own extension, built from
# /root/reference into oracle/_ref by oracle/build_ref.sh (falls back to the
# oracle port when that build is absent)
# ----------------------------------------------------------------------------
def import_cpu_impl():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "sequali", "_qc.abi3.so")):
        sys.path.insert(0, ref_dir)
        import sequali  # type: ignore
        return sequali, "reference"
    return None, "port"


def cpu_all_modules(text: bytes) -> float:
    """Seconds for one pass of the reference's hot loop over `text` (1 thread)."""
    ref, kind = import_cpu_impl()
    t0 = time.perf_counter()
    if ref is not None:
        mods = make_modules(ref)
        for arr in ref.FastqParser(io.BytesIO(text)):
            feed(mods, arr)
        mods["qc"].base_count_table()
    else:
        from tests import helpers as H
        H.oracle_single_end(text, ILLUMINA_ADAPTERS, chunk_records=384)
    return time.perf_counter() - t0


def cpu_baseline(data, reads: int):
    arr, n = data.to_host(reads)
    host = arr.tobytes()
    del arr
    dt = cpu_all_modules(host)
    _, kind = import_cpu_impl()
    return {"value": round(n * READ_LENGTH / dt / 1e9, 5), "unit": "Gbases/s", "cores": 1,
            "kind": kind, "reads_per_s": round(n / dt, 1),
            "sample": f"first {n} reads of the same synthetic input, uncompressed in RAM, default "
                      "128 KiB record arrays, one QC thread (the reference's QC loop is single-threaded)"}


def _ref_worker(args):
    path, seed, n, steps = args
    if path is not None:
        with open(path, "rb") as f:
            text = f.read()
    else:
        from sequali_b200 import synth
        text = synth.illumina_fastq(n, READ_LENGTH, seed=seed, n_tiles=936)
    return [cpu_all_modules(text) for _ in range(steps)]


_GEN_SAMPLE = r"""
import os, sys
sys.path.insert(0, {root!r})
from sequali_b200.device import DeviceFastq
import numpy as np
shards, n, total, out = {shards}, {n}, {total}, {out!r}
data = DeviceFastq.synth_illumina(shards * n, {length}, seed=2, chunk_reads=n, first_read=0, total_reads=total)
host, reads = data.to_host()
assert reads == shards * n
off = 0
for i, (_, nbytes, k) in enumerate(data.chunks):
    host[off:off + nbytes].tofile(os.path.join(out, "shard_%d.fastq" % i))
    off += nbytes
"""


def reference_sample_files(cores: int, n: int, total_reads: int):
    """The first cores x n reads of the SAME synthetic stream the CUDA arm measures (csrc/synth.cu, seed 2,
    tile runs of a `total_reads` stream), written by a helper process as one file per worker under
    /dev/shm: the processes that are timed load nothing but the reference.  None without a GPU."""
    import tempfile
    out = tempfile.mkdtemp(prefix="sq_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    code = _GEN_SAMPLE.format(root=ROOT, shards=cores, n=n, total=total_reads, out=out, length=READ_LENGTH)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    files = [os.path.join(out, f"shard_{i}.fastq") for i in range(cores)]
    if p.returncode != 0 or not all(os.path.exists(f) for f in files):
        return None, out
    return files, out


def run_reference(args):
    """The reference's CPU path on all host cores: one independent process per
    core over its own shard (the usage README.rst:160-166 recommends), each step
    a bounded sample of the workload: the first cores x ref_reads reads of the
    same device-generated stream the CUDA arm runs on."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import shutil
    cores = os.cpu_count() or 1
    n = args.ref_reads
    _, kind = import_cpu_impl()
    files, tmpdir = reference_sample_files(cores, n, args.reads)
    same_stream = files is not None
    jobs = [(files[i] if same_stream else None, 1000 + i, n, args.warmup + args.steps) for i in range(cores)]
    try:
        with mp.get_context("spawn").Pool(cores) as pool:
            t0 = time.perf_counter()
            times = pool.map(_ref_worker, jobs)
            total_wall = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    per_step = [max(t[args.warmup + s] for t in times) for s in range(args.steps)]
    dt = float(np.mean(per_step))
    value = cores * n * READ_LENGTH / dt / 1e9
    sample = (f"the first {cores} x {n} reads of the same synthetic stream as the CUDA arm (csrc/synth.cu, seed 2), "
              f"one contiguous shard per process" if same_stream else
              f"{cores} x {n} reads of the Python generator of the same recipe (no GPU for the device generator)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u64 (+f64 ordered sums)", "data": "synthetic",
        "config": {"workload": f"{args.reads} synthetic NovaSeq {READ_LENGTH} bp single-end reads per GPU, "
                               "936 tiles in runs, all default modules (QCMetrics, PerTileQuality, "
                               "OverrepresentedSequences, NanoStats, AdapterCounter, DedupEstimator)",
                   "reads_per_gpu": args.reads, "read_length": READ_LENGTH,
                   "sample": sample, "same_stream_as_cuda_arm": same_stream},
        "reads_per_s": round(cores * n / dt, 1),
        "cpu_baseline": {"value": round(value, 5), "unit": "Gbases/s", "cores": cores, "kind": kind,
                         "sample": f"{sample}; {cores} independent single-threaded processes (the reference's QC loop "
                                   f"has no internal threading); total wall {total_wall:.1f}s"},
        "e2e": {"value": round(value, 5), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU")
    ap.add_argument("--chunk-reads", type=int, default=1 << 22, help="reads per record array")
    ap.add_argument("--buffersize", type=int, default=64 << 20, help="e2e parser staging size")
    ap.add_argument("--e2e-reads", type=int, default=20_000_000)
    ap.add_argument("--e2e-window", type=int, default=512 << 20, help="bytes of host text per record array (e2e)")
    ap.add_argument("--cpu-reads", type=int, default=8_000_000)
    ap.add_argument("--ref-reads", type=int, default=500_000)
    ap.add_argument("--parity-reads", type=int, default=2_000_000, help="N > 1: reads of the sharded-vs-single check")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 leg")
    ap.add_argument("--bgzf-bytes", type=int, default=2 << 30, help="text bytes of the BGZF end-to-end leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly one JSON line: anything a native library writes to fd 1 meanwhile
    # (NCCL prints its version there when NCCL_DEBUG is set) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
