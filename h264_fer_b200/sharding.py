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
