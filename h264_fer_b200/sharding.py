"""Host-side work partitioning across GPUs (one process per GPU).

* Independent sequences (BASELINE config 5) shard with NO data-path collective: sequence s -> rank s % world.
* Macroblock-row bands of one picture (BASELINE config 4): contiguous bands, remainder rows to the first ranks (1080p:
  67 MB rows -> 9,9,9,8,8,8,8,8 on 8 GPUs). Phases A/S/C shard by band; the phase-B wavefront crosses the bands (the first MB
  row of a band needs the vectors of the band above). The exchange itself is not here: the kernels store the reconstruction and
  the last MB row's vectors straight into the peers' memory (bands.py, csrc/phase_b.cuh, phase_c.cuh); stage 2 reaches 279
  luma rows, more than a band, so every rank keeps the whole reference picture but waits, per picture, only for the bands within
  that halo (fh264_band_peers). Only the partitioning arithmetic and the small host-side reductions live here; they are what the
  gloo CPU tests exercise.
"""
from __future__ import annotations

from typing import List, Tuple


def sequences_for_rank(total: int, rank: int, world: int) -> List[int]:
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    return list(range(rank, total, world))


def mb_row_bands(mb_rows: int, world: int) -> List[Tuple[int, int]]:
    """[first_row, end_row) per rank; bands differ by at most one row, larger bands first."""
    if world < 1 or mb_rows < 0:
        raise ValueError("bad arguments")
    base, extra = divmod(mb_rows, world)
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        out.append((r, r + n))
        r += n
    return out


BAND_HALO = 304      # luma rows a band's search can reach (csrc/fh264_b200.cu FH_BAND_HALO: stage 2's 279 + block + box sums + 6-tap)
INDEX_TILE = 64


def band_wait_sets(bands: List[Tuple[int, int]], height: int) -> List[List[int]]:
    """For every rank the ranks it waits for at the picture barrier (host mirror of fh264_band_peers): those whose bands lie within
    its halo rows (rounded to the 64-row index tiles, + 16 rows of planes below and the 6-tap's 3), its direct neighbours (the rank
    below reads the band's last row of vectors), and — the relation is symmetric — every rank that waits for it."""
    n = len(bands)
    reads = []
    for a, (r0, r1) in enumerate(bands):
        top, bottom = r0 * 16, r1 * 16
        y0 = max(0, top - BAND_HALO) // INDEX_TILE * INDEX_TILE
        y1 = min(height, (bottom + BAND_HALO + INDEX_TILE - 1) // INDEX_TILE * INDEX_TILE)
        need0, need1 = y0 - 3, min(height, y1 + 16) + 3
        s = {r for r, (q0, q1) in enumerate(bands) if q0 * 16 < need1 and q1 * 16 > need0}
        s |= {a} | ({a + 1} if a + 1 < n else set()) | ({a - 1} if a > 0 else set())
        reads.append(s)
    return [sorted(reads[a] | {b for b in range(n) if a in reads[b]}) for a in range(n)]


def reduce_max(value: float) -> float:
    """Max over ranks of a per-rank device time (ms); identity without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local: List[int]) -> List[List[int]]:
    """All ranks' small integer lists (e.g. per-sequence picture counts) on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(local)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local))
    return out
