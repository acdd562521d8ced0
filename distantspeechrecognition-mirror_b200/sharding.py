"""Partition of independent recordings over ranks: recordings share no state
(`reset()` zeroes every node, reference modulated/modulated.cc:454-459, beamformer/beamformer.cc:1202-1212),
so the hot path has NO data-path collective; ranks only agree on (a) who owns which recording and
(b) the slowest rank's time, which is what throughput is quoted on."""
from __future__ import annotations


def owner_of(index: int, world: int) -> int:
    """Round-robin owner of recording `index`."""
    return index % world


def shard(n_recordings: int, rank: int, world: int) -> list:
    """Indices of the recordings rank `rank` processes."""
    return [i for i in range(n_recordings) if owner_of(i, world) == rank]


def greedy_by_length(lengths, world: int) -> list:
    """Length-balanced alternative: longest first onto the least loaded rank.  Returns owner per recording."""
    load = [0] * world
    owner = [0] * len(lengths)
    for i in sorted(range(len(lengths)), key=lambda k: -lengths[k]):
        r = min(range(world), key=lambda k: load[k])
        owner[i] = r
        load[r] += lengths[i]
    return owner


def job_throughput(units_per_rank, seconds_per_rank) -> float:
    """Whole-job rate = all units / the slowest rank's time."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)
