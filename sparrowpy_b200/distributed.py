"""Receiver-sharded energy exchange over several GPUs of one box.

north_star (3): receiver-side patches are sharded across 1/2/4/8 GPUs with an
all-gather of the per-order patch energy over NVLink.  One process per GPU
(``torch.distributed``); rank r owns the receiver patches
``[r*S, min((r+1)*S, N))``, ``S = ceil(N / world)`` rounded up to the receiver tile.
Every rank holds the full previous-order histogram ``E_{k-1}`` (all senders),
computes its slice of ``E_k`` (stage 1 + stage 2 restricted to its receivers) and the
slices are exchanged with an all-gather per order.  ``E_total`` stays sharded until
the end.

Bands never mix (RadiosityFast.py:1137-1143 acts per band), and the histogram layout
is band-major, so with more than one band the schedule is pipelined per band: band
b's slices are all-gathered on a side stream while the kernels of the other bands
run; order k+1 of band b only waits for band b's gather of order k.

The buffers are allocated for ``n_alloc = world * S`` patches per band so that every
shard has the same number of rows (NCCL all-gather needs equal counts); the padding
patches own no pairs and stay zero.

Communication modes (``SPB_COMM``):

``p2p`` (default) and ``multicast``: the ping-pong buffers are torch symmetric-memory allocations mapped into
every rank; stage 2 (``spb_exchange_mix_fused``) stores each ``E_k`` element of the
rank's receivers directly into all ranks' buffers -- one NVSwitch multicast store
(``multimem.st``) or one NVLink P2P store per peer -- so the all-gather is fused into
the compute kernel and only a cross-rank barrier separates the orders of a band.
``nccl``: local stores + ``all_gather_into_tensor`` on a side stream, pipelined
per band.
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from .exchange import EnergyHistogram, launch_gather

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

    ``compute(prev, cur, total, b_lo, b_hi)`` is the local step for a band range; the
    default launches the CUDA kernels through the C ABI.  Tests substitute a CPU
    function to exercise the collective choreography under gloo.
    """

    def __init__(self, tables, n_samples, device, group=None, compute=None,
                 layout=None, need_orders=True, gather_total=True, local=False):
        self.t = tables
        self.gather_total = gather_total     # False: E_total stays sharded (own rows only)
        self.group = group
        # local=True: a single-device run even when a process group is initialised (the
        # class API under torchrun must not silently turn into a collective)
        sharded = dist.is_initialized() and not local
        self.rank = dist.get_rank(group) if sharded else 0
        self.world = dist.get_world_size(group) if sharded else 1
        self.n_samples = n_samples
        self.j_lo, self.j_hi, self.shard = shard_range(tables.n_patches, self.rank,
                                                       self.world)
        self.n_alloc = self.shard * self.world if self.world > 1 else tables.n_patches
        if layout is None:
            layout = _lib.exchange_layout(n_samples, tables.max_delay, tables.dtype)
        self.t_pad, self.pad = layout
        self.ld = self.t_pad + self.pad
        tdt = _lib.torch_dtype(tables.dtype)
        t = tables
        rows = t.n_bands * self.n_alloc * t.n_dirs
        self.cuda = torch.device(device).type == "cuda"
        self.comm = "local"
        self.handles = None
        if self.world > 1:
            # P2P stores cost (world-1) x the shard in NVLink egress, a multicast store
            # 1 x (the switch replicates); measured on this pool (2 and 4 GPUs, C2) the
            # multimem.st path is nevertheless 10-20 % slower, so P2P is the default
            self.comm = os.environ.get("SPB_COMM", "p2p") if self.cuda else "gloo"
        if self.comm in ("multicast", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                self.e_a = symm_mem.empty((rows, self.ld), dtype=tdt, device=device)
                self.e_b = symm_mem.empty((rows, self.ld), dtype=tdt, device=device)
                self.handles = {self.e_a.data_ptr(): symm_mem.rendezvous(self.e_a, grp),
                                self.e_b.data_ptr(): symm_mem.rendezvous(self.e_b, grp)}
                if self.comm == "multicast" and not all(
                        h.multicast_ptr for h in self.handles.values()):
                    self.comm = "p2p"
                self.e_a.zero_()
                self.e_b.zero_()
            except Exception as exc:  # noqa: BLE001
                if os.environ.get("SPB_COMM"):
                    raise
                print(f"[sparrowpy_b200] symmetric memory unavailable ({exc}); "
                      "using NCCL all-gather", flush=True)
                self.comm, self.handles = "nccl", None
        if self.handles is None and need_orders:
            self.e_a = torch.zeros((rows, self.ld), dtype=tdt, device=device)
            self.e_b = torch.zeros((rows, self.ld), dtype=tdt, device=device)
        elif self.handles is None:
            self.e_a = self.e_b = None           # order 0 only: no ping-pong buffers
        self.e_total = torch.zeros((rows, self.ld), dtype=tdt, device=device)
        g_rows = t.n_bands * t.n_classes * t.n_patches if need_orders else 1
        self.g = torch.empty((max(g_rows, 1), self.ld), dtype=tdt, device=device)
        self.compute = compute or self._cuda_order
        self.comm_stream = torch.cuda.Stream(device=device) if (
            self.cuda and self.comm == "nccl") else None
        self.side_stream = torch.cuda.Stream(device=device) if (
            self.cuda and self.handles is not None) else None
        self.gather_fn = self._gather
        self.order_fn = self._order_fused

    # -- local kernels -------------------------------------------------------
    def _gather(self, prev, b_lo, b_hi, j_lo=None, j_hi=None):
        """Stage 1 for bands [b_lo, b_hi) and receivers [j_lo, j_hi) (default: the shard)."""
        j_lo = self.j_lo if j_lo is None else j_lo
        j_hi = self.j_hi if j_hi is None else j_hi
        launch_gather(self.t, prev, self.g, self.cta_order(j_lo, j_hi), self.n_alloc, b_lo,
                      b_hi, j_lo, j_hi, self.t_pad, self.ld, self.pad)

    def _cuda_order(self, prev, cur, total, b_lo, b_hi):
        if self.fused_order():
            self.order_fn(prev, cur, total, b_lo, b_hi)
            return
        self.gather_fn(prev, b_lo, b_hi)
        self._mix(cur, total, b_lo, b_hi)

    def fused_order(self):
        """True when one reflection order is ONE kernel: a single BRDF class and direction
        (diffuse walls) with the tensor-memory gather -- stage 2 and the delivery of E_k to
        every rank then happen in the gather's epilogue (spb_exchange_order_fused), tile by
        tile, overlapped with the tiles still running."""
        t = self.t
        return (self.cuda and t.win_recs is not None and t.n_classes == 1 and t.n_dirs == 1
                and self.comm != "multicast"
                and os.environ.get("SPB_FUSED_ORDER", "1") != "0"
                and os.environ.get("SPB_GATHER", "tmem") == "tmem")

    def _order_fused(self, prev, cur, total, b_lo, b_hi):
        import ctypes
        t = self.t
        if self.handles is None:
            dests = [cur.data_ptr()]
        else:
            dests = [int(p) for p in self.handles[cur.data_ptr()].buffer_ptrs]
        ptrs = (ctypes.c_uint64 * len(dests))(*dests)
        _lib.call("spb_exchange_order_fused", prev, ptrs, _lib.I32(len(dests)), total, t.coef,
                  t.win_ptr, t.win_recs, self.cta_order(), t.n_patches, self.n_alloc,
                  t.n_bands, b_lo, b_hi, self.j_lo, self.j_hi, self.t_pad, self.ld, self.pad,
                  t.win_w, _lib.I32(t.dtype), _lib.stream_ptr())

    def cta_order(self, j_lo=None, j_hi=None):
        """Launch order of the tiles of receivers [j_lo, j_hi) (default: the shard): longest
        record lists first (LPT), so the short tiles fill the tail of the grid.  (A class-major
        "neighbouring receivers together" order for L2 locality was measured on C2 and makes no
        difference: 13.5 ms either way, profiles/r02_tile_order_c2.txt.)"""
        j_lo = self.j_lo if j_lo is None else j_lo
        j_hi = self.j_hi if j_hi is None else j_hi
        cache = self.__dict__.setdefault("_cta_orders", {})
        if (j_lo, j_hi) not in cache:
            t = self.t
            n_r = SHARD_ALIGN
            n_blocks = -(-t.n_patches // n_r)
            jb_lo, jb_hi = j_lo // n_r, -(-j_hi // n_r)
            ptr = t.tile_ptr
            counts = (ptr[1:] - ptr[:-1]).view(t.n_classes, n_blocks)
            local = counts[:, jb_lo:jb_hi].reshape(-1)
            cache[(j_lo, j_hi)] = torch.argsort(local, descending=True, stable=True).to(
                torch.int32).contiguous()
        return cache[(j_lo, j_hi)]

    def _mix(self, cur, total, b_lo, b_hi, j_lo=None, j_hi=None):
        """Stage 2; with symmetric buffers it also delivers E_k to every rank."""
        import ctypes
        t = self.t
        code = _lib.I32(t.dtype)
        st = _lib.stream_ptr()
        j_lo = self.j_lo if j_lo is None else j_lo
        j_hi = self.j_hi if j_hi is None else j_hi
        if self.handles is None:
            _lib.call("spb_exchange_mix", self.g, cur, total, t.seg_ptr, t.coef, t.n_patches,
                      self.n_alloc, t.n_classes, t.n_dirs, t.n_bands, b_lo, b_hi, j_lo,
                      j_hi, self.t_pad, self.ld, self.pad, code, st)
            return
        hdl = self.handles[cur.data_ptr()]
        ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        mc = ctypes.c_void_p(int(hdl.multicast_ptr) if self.comm == "multicast" else 0)
        _lib.call("spb_exchange_mix_fused", self.g, ptrs, _lib.I32(self.world), mc, total,
                  t.seg_ptr, t.coef, t.n_patches, self.n_alloc, t.n_classes, t.n_dirs,
                  t.n_bands, b_lo, b_hi, j_lo, j_hi, self.t_pad, self.ld, self.pad,
                  code, st)

    def _barrier(self, channel=0):
        next(iter(self.handles.values())).barrier(channel)

    # -- row bookkeeping -----------------------------------------------------
    def band_rows(self, buf, b):
        """All rows of band ``b`` (every patch)."""
        n = self.n_alloc * self.t.n_dirs
        return buf[b * n:(b + 1) * n]

    def shard_rows(self, buf, b, rank=None):
        """Rows of band ``b`` owned by ``rank``."""
        rank = self.rank if rank is None else rank
        d = self.t.n_dirs
        r0 = (b * self.n_alloc + rank * self.shard) * d
        return buf[r0:r0 + self.shard * d]

    def _all_gather_band(self, buf, b):
        dist.all_gather_into_tensor(self.band_rows(buf, b), self.shard_rows(buf, b),
                                    group=self.group)

    # -- driver --------------------------------------------------------------
    def init(self, e0, delay0):
        """Initial energy (order 0): replicated into e_a (every rank needs all
        senders), own shard only into e_total.

        e0 (N, D, B) with delay0 (N,), or a batch of sources e0 (S, N, D, B) with
        delay0 (S, N) when the tables were tiled for S sources."""
        t = self.t
        if self.e_b is not None:
            self.e_b.zero_()
        if e0.dim() == 3:
            e0, delay0 = e0[None], delay0[None]
        n_src = e0.shape[0]
        nb_src = t.n_bands // n_src
        if self.cuda:
            tdt = _lib.torch_dtype(t.dtype)
            self.e_total.zero_()
            if self.e_a is not None:
                self.e_a.zero_()
            for s in range(n_src):
                _lib.call("spb_exchange_scatter", self.e_total, self.e_a,
                          t.to_internal(e0[s]).to(tdt).contiguous(),
                          t.to_internal(delay0[s]).to(torch.int32).contiguous(),
                          t.n_patches, self.n_alloc, t.n_dirs, nb_src, s * nb_src, t.n_bands,
                          self.n_samples, self.ld, self.pad, _lib.I32(t.dtype),
                          _lib.stream_ptr())
        else:  # CPU path of the gloo tests
            self.e_total.zero_()
            self.e_a.zero_()
            for s in range(n_src):
                e0s, d0s = t.to_internal(e0[s]), t.to_internal(delay0[s])
                n, d, nb = e0s.shape
                for b in range(nb):
                    for i in range(n):
                        dl = int(d0s[i])
                        if dl < self.n_samples:
                            r0 = ((s * nb_src + b) * self.n_alloc + i) * d
                            self.e_total[r0:r0 + d, self.pad + dl] += e0s[i, :, b].to(
                                self.e_total.dtype)
            self.e_a.copy_(self.e_total)
        if self.world > 1:
            keep = [self.shard_rows(self.e_total, b).clone() for b in range(t.n_bands)]
            self.e_total.zero_()
            for b in range(t.n_bands):
                self.shard_rows(self.e_total, b).copy_(keep[b])

    def run(self, max_order):
        t = self.t
        nb = t.n_bands
        prev, cur = self.e_a, self.e_b
        if self.world == 1:
            for _ in range(max_order):
                self.compute(prev, cur, self.e_total, 0, nb)
                prev, cur = cur, prev
        elif self.handles is not None:
            # fused exchange: stage 2 stores into every rank's buffer; a cross-rank
            # barrier per order separates "all peers wrote E_k" from "E_k is read".
            # The bands are split into two groups on two streams so that one group's
            # store-bound stage 2 (NVLink) overlaps the other group's stage 1 (SMs).
            main = torch.cuda.current_stream()
            groups = [(0, nb)] if nb < 2 else [(0, (nb + 1) // 2), ((nb + 1) // 2, nb)]
            streams = [main, self.side_stream][:len(groups)]
            self._barrier()                      # every rank finished init()
            for st in streams[1:]:
                st.wait_stream(main)
            for k in range(max_order):
                for gi, (b0, b1) in enumerate(groups):
                    with torch.cuda.stream(streams[gi]):
                        if k > 0:
                            self._barrier(gi)
                        self.compute(prev, cur, self.e_total, b0, b1)
                prev, cur = cur, prev
            for st in streams[1:]:
                main.wait_stream(st)
            self._barrier()                      # all stores landed before buffers are reused
            for b in range(nb if self.gather_total else 0):
                self._all_gather_band(self.e_total, b)
        elif not self.cuda:
            for _ in range(max_order):
                self.compute(prev, cur, self.e_total, 0, nb)
                for b in range(nb):
                    self._all_gather_band(cur, b)
                prev, cur = cur, prev
            for b in range(nb if self.gather_total else 0):
                self._all_gather_band(self.e_total, b)
        else:
            comp = torch.cuda.current_stream()
            comm = self.comm_stream
            comm.wait_stream(comp)
            ready = [None] * nb           # band b of `prev` is complete on every rank
            for _ in range(max_order):
                nxt = []
                for b in range(nb):
                    if ready[b] is not None:
                        comp.wait_event(ready[b])
                    self.compute(prev, cur, self.e_total, b, b + 1)
                    done = torch.cuda.Event()
                    done.record(comp)
                    comm.wait_event(done)
                    with torch.cuda.stream(comm):
                        self._all_gather_band(cur, b)
                        ev = torch.cuda.Event()
                        ev.record(comm)
                    nxt.append(ev)
                ready = nxt
                prev, cur = cur, prev
            for ev in ready:
                if ev is not None:
                    comp.wait_event(ev)
            for b in range(nb if self.gather_total else 0):
                self._all_gather_band(self.e_total, b)
        return EnergyHistogram(self.e_total, t.n_patches, t.n_dirs, t.n_bands,
                               self.n_samples, self.pad, n_alloc=self.n_alloc, tables=t)


class ShardedHistogram:
    """``sum_k E_k`` of a large scene, kept sharded: this rank holds the rows of its own
    receiver patches only, ``data[(band * shard + (p - j_lo)) * D + dir, PAD + t]`` with
    ``p`` the internal patch index.  Receiver collection sums the rank's patches and
    all-reduces the (R, B, T) result -- the full histogram never exists anywhere."""

    def __init__(self, data, tables, n_samples, pad, j_lo, j_hi, shard, group=None,
                 collect=None):
        self.data, self.tables = data, tables
        self.n_samples, self.pad = n_samples, pad
        self.j_lo, self.j_hi, self.shard = j_lo, j_hi, shard
        self.group = group
        self.collect = collect          # CPU stand-in for exchange.collect_mono (tests)

    def local(self):
        """The rank's rows as an EnergyHistogram over ``shard`` local patches."""
        t = self.tables
        return EnergyHistogram(self.data, self.shard, t.n_dirs, t.n_bands, self.n_samples,
                               self.pad, n_alloc=self.shard)

    def dense_local(self):
        """(j_hi - j_lo, D, B, T): the rank's receivers in internal order."""
        return self.local().dense()[:self.j_hi - self.j_lo]

    def collect_mono(self, rdir, shift, scale):
        """``collect_energy_receiver_mono`` over all ranks: rdir, shift (R, N) int32 and
        scale (R, N, B) in the caller's patch numbering, as for exchange.collect_mono."""
        from . import exchange
        t = self.tables
        own = slice(self.j_lo, self.j_lo + self.shard)

        def mine(x):
            x = t.to_internal(x, 1)
            if x.shape[1] < self.j_lo + self.shard:          # last shard: pad with zeros
                padn = self.j_lo + self.shard - x.shape[1]
                x = torch.cat([x, x.new_zeros((x.shape[0], padn) + x.shape[2:])], 1)
            return x[:, own].contiguous()

        fn = self.collect or exchange.collect_mono
        out = fn(self.local(), mine(rdir), mine(shift), mine(scale))
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(out, group=self.group)
        return out


class BandwiseExchange:
    """Memory-bounded schedule for large scenes (BASELINE config 5: ~100 k patches,
    16 directions, 8 bands).  Bands never mix (RadiosityFast.py:1137-1143 acts per band
    and the BRDF coefficients carry the band index), so the recursion is run band block
    by band block: the ping-pong buffers and the block's ``E_total`` exist for ONE block
    at a time (the inner :class:`ShardedExchange`, re-used), and after a block only the
    rows of this rank's receivers are kept.  Peak memory per rank is
    ``3 * block * N * D * LD + B * N / world * D * LD`` elements instead of
    ``3 * B * N * D * LD``."""

    def __init__(self, tables, n_samples, device, band_block=1, group=None, compute=None,
                 collect=None):
        import dataclasses
        self.t = tables
        self.n_samples = n_samples
        self.band_block = max(1, min(int(band_block), tables.n_bands))
        self.group = group
        self.collect = collect
        view = dataclasses.replace(
            tables, coef=tables.coef[:, :, :self.band_block].contiguous(),
            n_bands=self.band_block)
        self.sx = ShardedExchange(view, n_samples, device, group=group, compute=compute,
                                  gather_total=False)
        sx = self.sx
        # rows this rank keeps per band: its shard (all shards are padded to the same
        # size when world > 1), or simply every patch on a single GPU
        self.own = sx.shard if sx.world > 1 else tables.n_patches
        rows = tables.n_bands * self.own * tables.n_dirs
        self.total = torch.zeros((rows, sx.ld), dtype=sx.e_total.dtype, device=device)

    def _block_tables(self, b0, b1):
        import dataclasses
        return dataclasses.replace(self.t, coef=self.t.coef[:, :, b0:b1].contiguous(),
                                   n_bands=b1 - b0)

    def run(self, e0, delay0, max_order):
        """e0 (N, D, B), delay0 (N,) in the caller's numbering -> ShardedHistogram."""
        t, sx = self.t, self.sx
        d, nb = t.n_dirs, t.n_bands
        self.total.zero_()
        for b0 in range(0, nb, self.band_block):
            b1 = min(nb, b0 + self.band_block)
            if b1 - b0 != sx.t.n_bands:
                raise ValueError("n_bands must be a multiple of band_block")
            sx.t = self._block_tables(b0, b1)
            sx.init(e0[:, :, b0:b1].contiguous(), delay0)
            sx.run(max_order)
            for b in range(b0, b1):                  # keep the rank's rows of the block
                r0 = b * self.own * d
                src0 = ((b - b0) * sx.n_alloc + sx.j_lo) * d
                self.total[r0:r0 + self.own * d].copy_(
                    sx.e_total[src0:src0 + self.own * d])
        return ShardedHistogram(self.total, t, self.n_samples, sx.pad, sx.j_lo, sx.j_hi,
                                self.own, group=self.group, collect=self.collect)


# ---------------------------------------------------------------------------
# sharded bake: SURVEY.md 8(e) "pair tiles are independent"
# ---------------------------------------------------------------------------
def route_directed(dest, int_fields, ff, world, group=None):
    """Send every directed pair to the rank that owns its receiver: ``dest`` (M,) rank per
    entry, ``int_fields`` (M, K) int32, ``ff`` (M,) float64.  One variable-size all-to-all
    per tensor (the only collective of the bake).  Returns the entries this rank received."""
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world).to(torch.int64)
    ints, ff = int_fields[order].contiguous(), ff[order].contiguous()
    if world == 1:
        return ints, ff
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    send, recv = counts.tolist(), recv_counts.tolist()
    ints_out = ints.new_empty((sum(recv), ints.shape[1]))
    ff_out = ff.new_empty(sum(recv))
    dist.all_to_all_single(ints_out, ints, output_split_sizes=recv, input_split_sizes=send,
                           group=group)
    dist.all_to_all_single(ff_out, ff, output_split_sizes=recv, input_split_sizes=send,
                           group=group)
    return ints_out, ff_out


def sharded_bake_tables(rad, speed_of_sound, dt, n_samples, group=None, part=None,
                        n_parts=None, route=None):
    """``bake_geometry`` (RadiosityFast.py:369-433) + the exchange tables of THIS rank's
    receiver shard, with the bake itself split over the ranks: rank r evaluates the rows
    ``bake.triangle_row_range(N, r, world)`` of the visibility matrix (equal numbers of
    (i < j) entries), the form factors, distances and BRDF direction indices of the visible
    pairs it finds there, and sends each directed pair to the owner of its receiver.  No rank
    ever holds the (N, N) matrix or the pair list of the whole scene.  The tables are
    identical to ``rad._pair_tables(..., n_shards=world, shard=rank)`` after an unsharded
    bake (entries are sorted by (segment, sender) on arrival).

    ``part`` / ``n_parts`` / ``route`` let a single process play several ranks (tests)."""
    import numpy as np
    from . import bake, exchange, geometry
    sharded = dist.is_initialized() and part is None
    rank = dist.get_rank(group) if sharded else (part or 0)
    world = dist.get_world_size(group) if sharded else (n_parts or 1)
    g = rad._geom()
    dev = rad._device
    n = rad.n_patches
    lo, hi = bake.triangle_row_range(n, rank, world)
    vis_rows = bake.visibility_p2p_grouped(g["center"], g["normal"], g["points"],
                                           rad._patch_to_wall_ids, row_range=(lo, hi))
    pairs = bake.visible_pairs(vis_rows)
    pairs[:, 0] += lo
    del vis_rows
    ff, _ = bake.form_factors(g["points"], g["normal"], g["area"], pairs)
    with_brdf = rad._brdf_incoming_directions is not None
    if with_brdf:
        vi, vo, brdf, bidx = rad._brdf_tables()
        n_in, n_out = vi.shape[1], vo.shape[1]
        vi_d, vo_d = torch.from_numpy(vi).to(dev), torch.from_numpy(vo).to(dev)
    else:
        n_in = n_out = 1
        vi_d = vo_d = None
    dist_p, out_dir, in_dir = bake.pair_geometry(g["center"], g["wall_ids"], pairs, vi_d, vo_d)
    n_bins = 1 if rad._frequencies is None else rad.n_bins
    air = (np.zeros(n_bins) if rad._air_attenuation is None
           else np.real(rad._air_attenuation).astype(float))
    sender, receiver, ff_dir = exchange.directed_pairs(pairs, ff, g["area"])
    if with_brdf:
        coef = np.exp(-air)[None, None, :] * brdf.reshape(-1, n_out, n_bins)
        cls = torch.from_numpy(bidx).to(dev)[g["wall_ids"][sender]] * n_in + in_dir.long()
    else:
        coef = np.exp(-air)[None, None, :] * np.ones((1, 1, n_bins))
        cls = torch.zeros_like(sender)
    delay = bake.delay_bins(dist_p, speed_of_sound, dt)
    delay = torch.stack([delay, delay], dim=1).reshape(-1)
    n_pairs_local = int(pairs.shape[0])
    del pairs, ff, dist_p, in_dir
    perm, n_internal = geometry.compact_patch_order(rad._patches_points, rad._patch_to_wall_ids,
                                                    n_shards=world)
    perm_d = torch.from_numpy(perm).to(dev)
    j_lo, j_hi, size = shard_range(n_internal, rank, world)
    dest = (perm_d[receiver] // size).clamp_(max=world - 1)
    ints = torch.stack([sender.to(torch.int32), receiver.to(torch.int32),
                        out_dir.to(torch.int32), cls.to(torch.int32),
                        delay.to(torch.int32)], dim=1)
    kept = delay[delay < n_samples]          # pairs with delay >= T never contribute
    stats = torch.tensor([n_pairs_local, int(kept.max().item()) if kept.numel() else 0],
                         dtype=torch.int64, device=dev)
    del kept
    del sender, receiver, out_dir, cls, delay
    if route is not None:
        ints, ff_dir, stats = route(dest, ints, ff_dir, stats)
    else:
        ints, ff_dir = route_directed(dest, ints, ff_dir, world, group)
        if world > 1:
            n_tot = stats[:1].clone()
            dist.all_reduce(n_tot, group=group)
            d_max = stats[1:].clone()
            dist.all_reduce(d_max, op=dist.ReduceOp.MAX, group=group)
            stats = torch.cat([n_tot, d_max])
    n_pairs, max_delay = int(stats[0].item()), int(stats[1].item())
    tables = exchange.build_pair_tables(
        ints[:, 0].long(), ints[:, 1].long(), ff_dir, ints[:, 4].long(), ints[:, 2],
        ints[:, 3].long(), torch.from_numpy(np.ascontiguousarray(coef)).to(dev), n, n_samples,
        rad._dtype, rank=perm_d, n_internal=n_internal,
        receiver_range=(j_lo, j_hi) if world > 1 else None,
        max_delay=max_delay, n_directed=2 * n_pairs)
    return tables, n_pairs
