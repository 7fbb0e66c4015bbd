"""Multi-GPU sharding of independent blocks: contiguous block ranges per rank, no collective.

Blocks of an independent-block frame share nothing (FLG.BlockIndependence, header/write.go:37-39), so rank r
of W simply owns blocks [r*ceil(N/W), (r+1)*ceil(N/W)).  The only cross-shard quantity is frame ORDER:
record sizes are prefix-summed to place each shard's records and to produce the dstMark of every block
(what writeLoop's pending map does on one host, async/writer.go:316-348).
"""
from __future__ import annotations


def shard_blocks(nblk: int, world: int) -> list[tuple[int, int]]:
    """[b0, b1) per rank; contiguous, disjoint, covering, sizes differ by at most one chunk."""
    per = -(-nblk // world) if world > 0 else 0
    return [(min(r * per, nblk), min((r + 1) * per, nblk)) for r in range(world)]


def record_offsets(rec_len_per_rank: list[list[int]], header_size: int) -> list[list[int]]:
    """Frame offset (dstMark) of every block given each rank's record lengths, in rank order."""
    out, pos = [], header_size
    for lens in rec_len_per_rank:
        offs = []
        for n in lens:
            offs.append(pos)
            pos += n
        out.append(offs)
    return out
