"""sequali_b200.report -- the aggregation half of the reference's report (``report_modules.py``) on the
device-resident tables of the B200 collectors.

The reference's report first pulls whole tables to Python -- ``base_count_table()`` + ``phred_count_table()`` are
136 MB for 1 Mb reads -- and sums rows in Python loops (``aggregate_count_matrix`` :307-322,
``SequenceLengthDistribution.from_base_count_tables`` :575-637), and ``NanoStatsReport.from_nanostats`` (:1952-2046)
is a Python loop over every read.  The functions here return the same numbers (same names, same types) from
``QCMetrics.aggregate`` / ``NanoStats.report_tables`` (kernels in ``csrc/report.cu``); only the aggregated rows
leave the device.  Plotting / HTML stay where they are: a ``report_modules.py`` that wants the speed-up swaps
the bodies of the functions named below for these.  Results are plain dicts keyed by the
reference's dataclass field names.
"""
import math
from typing import Iterator, List, Sequence, Tuple

from ._qc import C, G, N, NUMBER_OF_NUCS, NUMBER_OF_PHREDS, NanoStats, QCMetrics

PERCENTILES = [1, 5, 10, 25, 50, 75, 90, 95, 99]


def equidistant_ranges(length: int, parts: int) -> Iterator[Tuple[int, int]]:
    """report_modules.py:258-269"""
    size, remainder = divmod(length, parts)
    start = 0
    for i in range(parts):
        part = size if i < parts - remainder else size + 1
        if part:
            yield start, start + part
            start += part


def logarithmic_ranges(length: int, min_distance: int = 5) -> Iterator[Tuple[int, int]]:
    """report_modules.py:272-290"""
    factor = 250_000_000 ** (1 / 400)
    i = start = 0
    while True:
        stop = round(factor ** i)
        i += 1
        if stop >= start + min_distance:
            yield start, stop
            start = stop
            if stop >= length:
                return


def data_ranges_for(max_length: int, graph_resolution: int = 200) -> List[Tuple[int, int]]:
    """calculate_stats, report_modules.py:2627-2631"""
    if max_length > 500:
        return list(logarithmic_ranges(max_length))
    return list(equidistant_ranges(max_length, graph_resolution))


def stringify_ranges(data_ranges) -> List[str]:
    """report_modules.py:293-297"""
    return [f"{start + 1}-{stop}" if start + 1 != stop else f"{start + 1}" for start, stop in data_ranges]


def qc_metrics_tables(metrics: QCMetrics, data_ranges: Sequence[Tuple[int, int]]) -> dict:
    """What qc_metrics_modules (:2537-2572) computes from the two big tables before it builds its modules:
    the aggregated matrices, the Summary numbers and SequenceLengthDistribution."""
    total_reads = metrics.number_of_reads
    thresholds = [int(p * total_reads / 100) for p in PERCENTILES]                # :597
    agg = metrics.aggregate(data_ranges, thresholds)
    base, phred = agg["base_matrix"], agg["phred_matrix"]
    n = len(data_ranges)
    summary_bases = [sum(base[c::NUMBER_OF_NUCS]) for c in range(NUMBER_OF_NUCS)] if n else [0] * NUMBER_OF_NUCS
    summary_phreds = [sum(phred[c::NUMBER_OF_PHREDS]) for c in range(NUMBER_OF_PHREDS)] if n else [0] * NUMBER_OF_PHREDS
    total_bases = sum(summary_bases)
    names = ["q1", "q5", "q10", "q25", "q50", "q75", "q90", "q95", "q99"]
    return {
        "aggregated_base_matrix": base,
        "aggregated_phred_matrix": phred,
        "x_labels": stringify_ranges(data_ranges),
        "summary": {"mean_length": total_bases / max(total_reads, 1), "minimum_length": agg["minimum_length"],
                    "maximum_length": metrics.max_length, "total_reads": total_reads, "total_bases": total_bases,
                    "q20_bases": sum(summary_phreds[5:]), "q20_reads": sum(metrics.phred_scores()[20:]),
                    "total_gc_bases": summary_bases[C] + summary_bases[G], "total_n_bases": summary_bases[N]},
        "sequence_length_distribution": {
            "length_ranges": ["0"] + stringify_ranges(data_ranges),
            "counts": [0] + agg["length_counts"],                                  # :596, 635 (entry 0 is never filled)
            **dict(zip(names, agg["threshold_lengths"])), "n50": agg["n50"], "n90": agg["n90"]},
    }


def _hour_minute(seconds: int) -> str:
    minutes = seconds // 60
    return f"{minutes // 60:02}:{minutes % 60:02}"


def nanostats_report(nanostats: NanoStats) -> dict:
    """NanoStatsReport.from_nanostats (:1952-2046), fields of the reference's dataclass."""
    if nanostats.skipped_reason:
        return {"x_labels": [], "time_bases": [], "time_reads": [], "time_active_channels": [],
                "qual_percentages_over_time": [], "per_channel_bases": {}, "per_channel_quality": {},
                "translocation_speed": [], "reads_with_parent": None, "total_reads": None,
                "skipped_reason": nanostats.skipped_reason}
    run_start_time = nanostats.minimum_time
    duration = nanostats.maximum_time - run_start_time
    time_per_slot = duration / 200
    time_interval = max(((math.ceil(time_per_slot) + 59) // 60) * 60, 1)
    time_ranges = [(start, start + time_interval) for start in range(0, duration + 1, time_interval)]
    t = nanostats.report_tables(run_start_time, time_interval, len(time_ranges))
    per_channel_bases = dict(zip(t["channels"], t["channel_bases"]))             # sorted by channel already
    per_channel_quality = {}
    for channel, bases, error in zip(t["channels"], t["channel_bases"], t["channel_cumulative_error"]):
        per_channel_quality[channel] = -10 * math.log10(error / bases) if bases else 0
    qual_percentages = [[] for _ in range(12)]
    for quals in t["time_qualities"]:
        total = sum(quals)
        for i, q in enumerate(quals):
            qual_percentages[i].append(q / max(total, 1))
    return {"x_labels": [f"{_hour_minute(a)}-{_hour_minute(b)}" for a, b in time_ranges],
            "qual_percentages_over_time": qual_percentages, "time_active_channels": t["time_active_channels"],
            "time_bases": t["time_bases"], "time_reads": t["time_reads"], "per_channel_bases": per_channel_bases,
            "per_channel_quality": per_channel_quality, "translocation_speed": t["translocation_speed"],
            "skipped_reason": nanostats.skipped_reason, "total_reads": nanostats.number_of_reads,
            "reads_with_parent": t["reads_with_parent"] if t["reads_with_parent"] > 0 else None}
