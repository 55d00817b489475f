"""Receiver-sharded energy exchange over several GPUs of one box.

north_star (3): receiver-side patches are sharded across 1/2/4/8 GPUs with an
all-gather of the per-order patch energy over NVLink.  One process per GPU
(``torch.distributed``); rank r owns the receiver patches
``[r*S, min((r+1)*S, N))`` with ``S = ceil(N / world)``.  Every rank holds the full
previous-order histogram ``E_{k-1}`` (all senders), computes its slice of ``E_k``
(stage 1 + stage 2 restricted to its receivers) and the slices are exchanged with
one all-gather per order.  ``E_total`` stays sharded until the end.

The histogram buffers are allocated for ``world * S`` patches so that every shard
has the same number of rows (NCCL all-gather needs equal counts); the padding
patches own no pairs and stay zero.
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from .exchange import EnergyHistogram


SHARD_ALIGN = 8


def shard_range(n_patches, rank, world):
    """Receiver patches owned by ``rank``: (j_lo, j_hi, shard_size)."""
    size = -(-n_patches // world)
    size = -(-size // SHARD_ALIGN) * SHARD_ALIGN     # tiles of 8 receivers stay whole
    lo = min(n_patches, rank * size)
    hi = min(n_patches, lo + size)
    return lo, hi, size


class ShardedExchange:
    """Per-order orchestration: local gather+mix, then all-gather of ``E_k``.

    ``compute`` is the per-order local step; the default launches the CUDA kernels
    through the C ABI.  Tests substitute a CPU function to exercise the collective
    choreography under gloo.
    """

    def __init__(self, tables, n_samples, device, group=None, compute=None,
                 layout=None):
        self.t = tables
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_samples = n_samples
        self.j_lo, self.j_hi, self.shard = shard_range(tables.n_patches, self.rank,
                                                       self.world)
        self.n_pad = self.shard * self.world
        if layout is None:
            layout = _lib.exchange_layout(n_samples, tables.max_delay, tables.dtype)
        self.t_pad, self.pad = layout
        self.ld = self.t_pad + self.pad
        tdt = _lib.torch_dtype(tables.dtype)
        self.db = tables.n_dirs * tables.n_bands
        rows = self.n_pad * self.db
        self.e_a = torch.zeros((rows, self.ld), dtype=tdt, device=device)
        self.e_b = torch.zeros((rows, self.ld), dtype=tdt, device=device)
        self.e_total = torch.zeros((rows, self.ld), dtype=tdt, device=device)
        g_rows = tables.n_classes * tables.n_patches * tables.n_bands
        self.g = torch.empty((max(g_rows, 1), self.ld), dtype=tdt, device=device)
        self.compute = compute or self._cuda_order

    # -- local kernels -------------------------------------------------------
    def _cuda_order(self, prev, cur, total):
        t = self.t
        code = _lib.I32(t.dtype)
        st = _lib.stream_ptr()
        if t.recs is not None and os.environ.get("SPB_GATHER", "tma") != "csr":
            _lib.call("spb_exchange_gather_tiled", prev, self.g, t.ent_ptr, t.recs,
                      t.n_patches, t.n_classes, t.n_bands, self.j_lo, self.j_hi, self.t_pad,
                      self.ld, self.pad, code, st)
        else:
            _lib.call("spb_exchange_gather", prev, self.g, t.seg_ptr, t.src, t.wgt, t.dly,
                      t.n_patches, t.n_classes, t.n_bands, self.j_lo, self.j_hi,
                      self.t_pad, self.ld, self.pad, code, st)
        _lib.call("spb_exchange_mix", self.g, cur, total, t.seg_ptr, t.coef, t.n_patches,
                  t.n_classes, t.n_dirs, t.n_bands, self.j_lo, self.j_hi, self.t_pad,
                  self.ld, self.pad, code, st)

    def _shard_rows(self, buf, rank=None):
        rank = self.rank if rank is None else rank
        r0 = rank * self.shard * self.db
        return buf[r0:r0 + self.shard * self.db]

    def _all_gather(self, buf):
        if self.world > 1:
            dist.all_gather_into_tensor(buf, self._shard_rows(buf), group=self.group)

    # -- driver --------------------------------------------------------------
    def init(self, e0, delay0):
        """Initial energy (order 0) into e_total and e_a on every rank (replicated:
        N*D*B scatters)."""
        t = self.t
        self.e_total.zero_()
        self.e_a.zero_()
        self.e_b.zero_()
        if e0.is_cuda:
            tdt = _lib.torch_dtype(t.dtype)
            n_rows = t.n_patches * self.db
            _lib.call("spb_exchange_init", self.e_total[:n_rows], self.e_a[:n_rows],
                      e0.to(tdt).contiguous(), delay0.to(torch.int32).contiguous(),
                      t.n_patches, self.db, self.n_samples, self.ld, self.pad,
                      _lib.I32(t.dtype), _lib.stream_ptr())
        else:  # CPU path of the gloo tests
            rows = torch.arange(t.n_patches * self.db)
            d = delay0.long().repeat_interleave(self.db)
            ok = d < self.n_samples
            vals = e0.reshape(-1).to(self.e_total.dtype)
            self.e_total[rows[ok], self.pad + d[ok]] += vals[ok]
            self.e_a[rows[ok], self.pad + d[ok]] += vals[ok]
        # e_total keeps only this rank's shard (the final all-gather assembles it)
        keep = self._shard_rows(self.e_total).clone()
        self.e_total.zero_()
        self._shard_rows(self.e_total).copy_(keep)

    def run(self, max_order):
        prev, cur = self.e_a, self.e_b
        for _ in range(max_order):
            self.compute(prev, cur, self.e_total)
            self._all_gather(cur)
            prev, cur = cur, prev
        self._all_gather(self.e_total)
        return EnergyHistogram(self.e_total[:self.t.n_patches * self.db], self.t.n_patches,
                               self.t.n_dirs, self.t.n_bands, self.n_samples, self.pad)
