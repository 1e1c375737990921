"""Band mode (BASELINE config 4): one picture coded by all GPUs of a node, split into macroblock-row bands.

One process per GPU (``torch.distributed`` only carries the setup: the CUDA-IPC handles of each rank's buffers). After
setup the data path is entirely on the devices: the phase-B wavefront crosses GPUs through progress flags that each rank
mirrors into the next rank's memory, the reconstructed bands travel as NVLink peer stores issued by phase C itself, and a
device-side barrier separates pictures — per halo: a rank only waits for the ranks whose bands its search can reach (304 luma rows),
so pictures are pipelined across the GPUs. Every rank must encode the same pictures in the same order; the whole reconstruction
is on every rank once ALL ranks have finished the picture (synchronise them on the host before ``download_recon``).
"""
from __future__ import annotations

import numpy as np

from . import sharding
from .native import MB_RESULT_DTYPE, Session


class BandSession:
    def __init__(self, width, height, device, rank=None, world=None, pipelined=True, gather=False):
        import torch.distributed as dist
        self.rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
        self.s = Session(width, height, batch=1, device=device)
        self.bands = sharding.mb_row_bands(height >> 4, self.world)
        self.row0, self.row1 = self.bands[self.rank]
        if self.row1 <= self.row0:
            raise ValueError("more ranks than macroblock rows")
        self.wmb = width >> 4
        self.s.band_config(self.rank, self.world, self.row0, self.row1)
        if self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, self.s.ipc_export(0))
            for r, blob in enumerate(blobs):
                if r != self.rank:
                    self.s.ipc_import(0, r, blob)
            if pipelined:
                # every rank knows every band: the picture barrier becomes a wait for the ranks within the halo (fh264_band_peers)
                self.s.band_peers(self.bands)
            if gather:
                # rank 0 collects the whole picture's records (phase C of every rank stores them there): device CAVLC of the slice
                self.s.band_gather(True)
            dist.barrier()

    @property
    def mb_slice(self):
        return slice(self.row0 * self.wmb, self.row1 * self.wmb)

    def upload_recon(self, y, cb, cr):
        self.s.upload_recon(0, y, cb, cr)

    def upload_source(self, y, cb, cr):
        self.s.upload_source(0, y, cb, cr)

    def encode_p(self, qp, window, maxdiff_set, basic=0, out=None, sync=True):
        """Records of THIS rank's band (a view of the band slice of a full-picture array)."""
        full = self.s.encode_p(qp, window, maxdiff_set, basic, out=out, sync=sync)
        return full[0][self.mb_slice]

    def gather_records(self, band_records):
        """All ranks' band records concatenated in raster order, on every rank (host-side, tests / bitstream writer)."""
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, np.asarray(band_records).tobytes())
        return np.concatenate([np.frombuffer(p, dtype=MB_RESULT_DTYPE) for p in parts])

    def download_recon(self):
        return self.s.download_recon(0)

    def close(self):
        import torch.distributed as dist
        self.s.sync()
        if self.world > 1:
            try:
                dist.barrier()          # nobody unmaps buffers a peer may still write
            except Exception:
                pass
        self.s.close()
