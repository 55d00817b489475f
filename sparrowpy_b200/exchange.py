"""Host side of the energy exchange and receiver collection.

Mirrors the array-level operators of the reference
(``_energy_exchange_init_energy``, ``_energy_exchange``, ``_collect_receiver_energy``;
reference RadiosityFast.py:1037-1185) on top of the C ABI in
include/sparrow_b200.h.  PyTorch is used for device memory, streams and index
bookkeeping only; all arithmetic on the energy histograms runs in the CUDA
kernels of csrc/exchange.cu.
"""
import os
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class PairTables:
    """Factored ``form_factors_tilde`` in receiver-major CSR form (device)."""
    seg_ptr: torch.Tensor      # (C*N + 1,) int64
    src: torch.Tensor          # (nnz,) int32   sender patch * D + outgoing dir
    wgt: torch.Tensor          # (nnz,) dtype   directed form factor
    dly: torch.Tensor          # (nnz,) int32   delay bins
    coef: torch.Tensor         # (C, D, B) dtype  exp(-air) * brdf
    n_patches: int
    n_classes: int
    n_dirs: int
    n_bands: int
    max_delay: int
    n_directed: int            # directed pairs before dropping delay >= T
    dtype: int
    ent_ptr: torch.Tensor = None   # (C*ceil(N/R) + 1,) int64  tiled records (TMA path)
    recs: torch.Tensor = None      # (n_records, record_bytes) uint8
    n_records: int = 0
    rank: torch.Tensor = None      # (n_user,) int64 caller's patch id -> internal index
    n_user: int = 0                # caller's patch count (n_patches is the internal one)
    win_ptr: torch.Tensor = None   # (C*ceil(N/R) + 1,) int64  window records (tmem gather)
    win_recs: torch.Tensor = None  # (n_records, 80) uint8
    win_w: int = 0                 # delay window of the records (4 or 10 bins)

    @property
    def tile_ptr(self):
        """Record offsets of whichever tiled gather these tables were built for."""
        return self.win_ptr if self.win_recs is not None else self.ent_ptr

    def tiled(self, n_sources):
        """Tables for a batch of ``n_sources`` sources: a source is one more group of
        independent channels, so source s simply occupies the bands [s*B, (s+1)*B)
        (same pair structure, BRDF coefficients repeated)."""
        import dataclasses
        if n_sources == 1:
            return self
        return dataclasses.replace(
            self, coef=self.coef.repeat(1, 1, n_sources).contiguous(),
            n_bands=self.n_bands * n_sources)

    def to_internal(self, per_patch, dim=0):
        """Reorder a per-patch tensor (patch axis ``dim``) into the internal patch
        numbering; holes of the internal numbering are zero."""
        if self.rank is None:
            return per_patch
        shape = list(per_patch.shape)
        shape[dim] = self.n_patches
        out = torch.zeros(shape, dtype=per_patch.dtype, device=per_patch.device)
        return out.index_copy_(dim, self.rank, per_patch)


def directed_pairs(pairs, ff_pairs, areas):
    """Expand visible pairs (lo < hi) into directed pairs.

    Entry 2p is lo->hi, entry 2p+1 is hi->lo -- the order in which the reference
    loop visits them (RadiosityFast.py:1124-1131).  The reverse form factor is
    ``ff * A_receiver / A_sender`` (RadiosityFast.py:1255-1256).
    """
    lo, hi = pairs[:, 0].long(), pairs[:, 1].long()
    sender = torch.stack([lo, hi], dim=1).reshape(-1)
    receiver = torch.stack([hi, lo], dim=1).reshape(-1)
    ff_rev = ff_pairs * areas[lo] / areas[hi]
    ff = torch.stack([ff_pairs, ff_rev], dim=1).reshape(-1)
    return sender, receiver, ff


def gather_kind(code=None):
    """Stage-1 kernel selected by ``SPB_GATHER``: ``tmem`` (default; window records, operands
    from tensor memory, exchange_tmem.cu; FP64 only), ``tma`` (bucket records, operands from
    shared memory, exchange_tma.cu) or ``csr`` (cross-check kernel).  FP32 tables fall back
    to ``tma``."""
    kind = os.environ.get("SPB_GATHER", DEFAULT_GATHER)
    if kind not in ("tmem", "tma", "csr"):
        raise ValueError(f"SPB_GATHER={kind!r}: use tmem, tma or csr")
    if kind == "tmem" and code is not None and code != _lib.F64:
        kind = "tma"
    return kind


DEFAULT_GATHER = "tmem"


def launch_gather(tables, prev, g, cta_order, n_alloc, b_lo, b_hi, j_lo, j_hi, t_pad, ld,
                  pad, kind=None):
    """Stage 1 of one order (``G`` from ``E_{k-1}``) with the kernel the tables were
    built for: receivers [j_lo, j_hi), bands [b_lo, b_hi)."""
    t = tables
    code, st = _lib.I32(t.dtype), _lib.stream_ptr()
    kind = kind or gather_kind(t.dtype)
    if t.win_recs is not None and kind != "csr":
        _lib.call("spb_exchange_gather_tmem", prev, g, t.win_ptr, t.win_recs, cta_order,
                  t.n_patches, n_alloc, t.n_classes, t.n_dirs, t.n_bands, b_lo, b_hi,
                  j_lo, j_hi, t_pad, ld, pad, t.win_w, code, st)
    elif t.recs is not None and kind != "csr":
        _lib.call("spb_exchange_gather_tiled", prev, g, t.ent_ptr, t.recs, cta_order,
                  t.n_patches, n_alloc, t.n_classes, t.n_dirs, t.n_bands, b_lo, b_hi, j_lo,
                  j_hi, t_pad, ld, pad, code, st)
    else:
        _lib.call("spb_exchange_gather", prev, g, t.seg_ptr, t.src, t.wgt, t.dly, t.n_patches,
                  n_alloc, t.n_classes, t.n_dirs, t.n_bands, b_lo, b_hi, j_lo, j_hi, t_pad,
                  ld, pad, code, st)


def build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n_patches,
                      n_samples, dtype, rank=None, n_internal=None, gather=None,
                      receiver_range=None, max_delay=None, n_directed=None):
    """Sort directed pairs into segments (class, receiver) and drop pairs whose
    delay is >= n_samples (they contribute nothing, RadiosityFast.py:1137-1140).

    ``rank`` (optional, (N,) int64) renumbers the patches inside the exchange:
    patch p becomes internal index ``rank[p] < n_internal`` (holes allowed).  The
    exchange is invariant under a relabelling of patches; a spatially compact
    numbering makes the 8 receivers of a tile close neighbours (more shared senders,
    directions and delay bins per record).  Histograms come back in the caller's
    numbering (EnergyHistogram.dense).

    ``receiver_range`` (optional, ``(lo, hi)`` in the internal numbering) keeps only the
    pairs whose receiver lies in that range -- the tables of one receiver shard of a
    multi-GPU run (distributed.shard_range).  The segment / tile index spaces and
    ``max_delay`` / ``n_directed`` stay those of the whole scene, so that every rank
    derives the same buffer layout.  ``max_delay`` / ``n_directed`` override those two
    quantities when the pair list handed in is already one shard's (sharded bake).
    """
    code = _lib.dtype_code(dtype)
    tdt = _lib.torch_dtype(code)
    n_classes, n_dirs, n_bands = coef.shape
    n_directed = int(sender.numel()) if n_directed is None else int(n_directed)
    keep = delay < n_samples
    sender, receiver, ff = sender[keep], receiver[keep], ff[keep]
    delay, out_dir, cls = delay[keep], out_dir[keep], cls[keep]
    n_user = int(n_patches)
    if rank is not None:
        rank = rank.to(sender.device).long().contiguous()
        n_patches = int(n_internal)
        sender, receiver = rank[sender.long()], rank[receiver.long()]
    if max_delay is None:
        max_delay = int(delay.max().item()) if delay.numel() else 0
    if receiver_range is not None:
        lo, hi = receiver_range
        own = (receiver >= lo) & (receiver < hi)
        sender, receiver, ff = sender[own], receiver[own], ff[own]
        delay, out_dir, cls = delay[own], out_dir[own], cls[own]
    seg = cls.long() * n_patches + receiver.long()
    key = seg * n_patches + sender.long()
    order = torch.argsort(key)
    seg = seg[order]
    counts = torch.bincount(seg, minlength=n_classes * n_patches)
    seg_ptr = torch.zeros(n_classes * n_patches + 1, dtype=torch.int64,
                          device=sender.device)
    seg_ptr[1:] = torch.cumsum(counts, 0)
    src = (sender[order] * n_dirs + out_dir[order].long()).to(torch.int32)
    gather = gather or gather_kind(code)
    ent_ptr = recs = win_ptr = win_recs = None
    win_w = 0
    if gather == "tmem" and code == _lib.F64:
        win_ptr, win_recs, win_w = build_window_records(
            sender, receiver, ff, delay, out_dir, cls, n_patches, n_dirs, n_classes, code)
        win_ptr, win_recs = device_window_records(win_ptr, win_recs)
    else:
        ent_ptr, recs = build_tile_records(sender, receiver, ff, delay, out_dir, cls,
                                           n_patches, n_dirs, n_classes, code)
    return PairTables(
        rank=rank, n_user=n_user, win_ptr=win_ptr, win_recs=win_recs, win_w=win_w,
        ent_ptr=ent_ptr, recs=recs,
        n_records=int((recs if recs is not None else win_recs).shape[0]),
        seg_ptr=seg_ptr.contiguous(), src=src.contiguous(),
        wgt=ff[order].to(tdt).contiguous(),
        dly=delay[order].to(torch.int32).contiguous(),
        coef=coef.to(tdt).contiguous(), n_patches=int(n_patches),
        n_classes=int(n_classes), n_dirs=int(n_dirs), n_bands=int(n_bands),
        max_delay=max_delay, n_directed=n_directed, dtype=code)


def tile_geometry(code):
    """(receivers per tile, delay bucket, record bytes) of the tiled gather kernel."""
    import ctypes
    lib = _lib.load()
    r, q, nbytes = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.spb_tile_geometry(ctypes.c_int(code), ctypes.byref(r), ctypes.byref(q),
                               ctypes.byref(nbytes))
    if rc != 0:
        raise _lib.SparrowB200Error(lib.spb_last_error().decode())
    return r.value, q.value, nbytes.value


def build_tile_records(sender, receiver, ff, delay, out_dir, cls, n_patches, n_dirs,
                       n_classes, code):
    """Union-of-senders records for the TMA-staged gather (csrc/exchange_tma.cu).

    A tile is R neighbouring receivers of one class; directed pairs that share the
    tile, the sender row and the 32-bin delay bucket are merged into one record
    ``{w[R], rel[R], src, dmin}``.  Pure index bookkeeping (sort / unique / scatter).
    """
    n_r, bucket, rec_bytes = tile_geometry(code)
    tdt = _lib.torch_dtype(code)
    dev = sender.device
    n_blocks = -(-n_patches // n_r)
    n_tiles = n_classes * n_blocks
    if sender.numel() == 0:
        return (torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev),
                torch.zeros((0, rec_bytes), dtype=torch.uint8, device=dev))
    n_rows = n_patches * n_dirs
    tile = cls.long() * n_blocks + receiver.long() // n_r
    slot = receiver.long() % n_r
    srow = sender.long() * n_dirs + out_dir.long()
    bkt = delay.long() // bucket
    n_bkt = int(bkt.max().item()) + 1
    key = (tile * n_rows + srow) * n_bkt + bkt
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    n_rec = uniq.numel()
    w = torch.zeros((n_rec, n_r), dtype=tdt, device=dev)
    w[inv, slot] = ff.to(tdt)
    # shift (delay - dmin) per slot; -1 marks an empty slot for the passes below
    shift = torch.full((n_rec, n_r), -1, dtype=torch.int16, device=dev)
    shift[inv, slot] = (delay.long() - bkt * bucket).to(torch.int16)
    # reload mask: bit s set when slot s is occupied and its shift differs from the
    # shift the kernel has loaded (that of the previous occupied slot); empty slots
    # repeat the previous shift so that they never trigger a reload
    mask = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    cur = torch.full((n_rec,), -1, dtype=torch.int16, device=dev)
    rel = torch.zeros((n_rec, n_r), dtype=torch.uint8, device=dev)
    for s_ in range(n_r):
        col = shift[:, s_]
        present = col >= 0
        change = present & (col != cur)
        mask |= change.long() << s_
        cur = torch.where(present, col, cur)
        rel[:, s_] = torch.clamp(cur, min=0).to(torch.uint8)
    dmin = (uniq % n_bkt) * bucket
    assert int(dmin.max().item()) < (1 << 24)
    packed = dmin | (mask << 24)
    packed = torch.where(packed >= (1 << 31), packed - (1 << 32), packed)   # as int32 bits
    meta = torch.stack([(uniq // n_bkt) % n_rows, packed], dim=1).to(torch.int32)
    recs = torch.cat([w.view(torch.uint8), rel, meta.contiguous().view(torch.uint8)],
                     dim=1).contiguous()
    assert recs.shape[1] == rec_bytes, (recs.shape, rec_bytes)
    counts = torch.bincount(uniq // (n_bkt * n_rows), minlength=n_tiles)
    ent_ptr = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
    ent_ptr[1:] = torch.cumsum(counts, 0)
    return ent_ptr.contiguous(), recs


def window_geometry(code):
    """(receivers per tile, widest delay window, record bytes) of the tensor-memory
    gather kernel (csrc/exchange_tmem.cu)."""
    import ctypes
    lib = _lib.load()
    r, q, nbytes = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.spb_window_geometry(ctypes.c_int(code), ctypes.byref(r), ctypes.byref(q),
                                 ctypes.byref(nbytes))
    if rc != 0:
        raise _lib.SparrowB200Error(lib.spb_last_error().decode())
    return r.value, q.value, nbytes.value


WINDOW_CHOICES = (4, 10)     # instantiations of k_gather_tmem


def _window_cover(delay_sorted, group, pos, n_groups, n_r, width, align=2):
    """Greedy cover of every group's sorted delays by windows [base, base + width] with
    ``(base + width) % align == 0`` (``align`` = 2: even base, the staged row starts on
    a 16-byte boundary; 4: on a 32-byte sector boundary, at the price of a window that
    starts up to 3 bins before the first delay).  Returns (starts_new_record (bool),
    base per entry)."""
    dev = delay_sorted.device
    base = torch.zeros(n_groups, dtype=torch.int64, device=dev)
    start = torch.zeros(delay_sorted.numel(), dtype=torch.bool, device=dev)
    ebase = torch.zeros(delay_sorted.numel(), dtype=torch.int64, device=dev)
    for p in range(n_r):
        idx = torch.nonzero(pos == p).reshape(-1)
        if idx.numel() == 0:
            break
        gp, dp = group[idx], delay_sorted[idx]
        new = dp > base[gp] + width if p else torch.ones_like(dp, dtype=torch.bool)
        nb = dp[new] - (dp[new] + width) % align
        base[gp[new]] = torch.where(nb < 0, dp[new] & ~1, nb)
        start[idx] = new
        ebase[idx] = base[gp]
    return start, ebase


def tmem_batch():
    """Records per pipeline hand-over of the tensor-memory gather (csrc/exchange_tmem.cu):
    every tile's record list is padded to a multiple of it."""
    return int(_lib.load().spb_tmem_batch())


def pad_record_lists(ent_ptr, recs, multiple):
    """Pad every tile's record list to a multiple of ``multiple`` with null records
    (w = 0, rel = 255, src = 0, dbase = 0: they add nothing)."""
    if multiple <= 1 or recs.shape[0] == 0:
        return ent_ptr, recs
    counts = ent_ptr[1:] - ent_ptr[:-1]
    padded = -(-counts // multiple) * multiple
    new_ptr = torch.zeros_like(ent_ptr)
    new_ptr[1:] = torch.cumsum(padded, 0)
    n_new = int(new_ptr[-1].item())
    out = torch.zeros((n_new, recs.shape[1]), dtype=torch.uint8, device=recs.device)
    out[:, 64:72] = 255                                     # rel of a null record
    tile = torch.repeat_interleave(torch.arange(counts.numel(), device=recs.device), counts)
    dst = torch.arange(recs.shape[0], device=recs.device) + (new_ptr[:-1] - ent_ptr[:-1])[tile]
    out[dst] = recs
    return new_ptr.contiguous(), out


def device_window_records(ent_ptr, recs):
    """Window records as the kernel reads them: every tile's list padded to the pipeline
    batch, and the shift bytes turned into TMEM column offsets -- ``2 * (delay - dbase)``, and
    0 for an empty slot, whose weight is 0 (the consumers read the records straight from
    shared memory, no per-record massaging on the device)."""
    ent_ptr, recs = pad_record_lists(ent_ptr, recs, tmem_batch())
    rel = recs[:, 64:72]
    recs[:, 64:72] = torch.where(rel == 255, torch.zeros_like(rel), rel * 2)
    return ent_ptr, recs


def build_window_records(sender, receiver, ff, delay, out_dir, cls, n_patches, n_dirs,
                         n_classes, code, width=None, align=None):
    """Records of the tensor-memory gather (csrc/exchange_tmem.cu).

    A tile is R neighbouring receivers of one class.  The directed pairs of one
    (tile, sender row) are covered greedily, in order of delay, by windows
    ``[dbase, dbase + W]`` with ``dbase`` even; each window is one record
    ``{w[R] f64, rel[R] u8 (delay - dbase, 255 = no pair), src i32, dbase i32}``.
    ``W`` is the narrowest instantiated window that costs at most 2 % more records than
    the widest one.  ``align`` = 2 (default) keeps ``dbase`` even -- the kernel stages rows in
    16-byte chunks; 4 is accepted for experiments.  Pure index bookkeeping (sort / cumsum /
    scatter).  Returns ``(ent_ptr, recs, W)``.
    """
    if align is None:
        align = 2
    if align not in (2, 4):
        raise ValueError("window alignment must be 2 or 4 bins")
    n_r, max_w, rec_bytes = window_geometry(code)
    tdt = _lib.torch_dtype(code)
    dev = sender.device
    n_blocks = -(-n_patches // n_r)
    n_tiles = n_classes * n_blocks
    if sender.numel() == 0:
        return (torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev),
                torch.zeros((0, rec_bytes), dtype=torch.uint8, device=dev), WINDOW_CHOICES[0])
    n_rows = n_patches * n_dirs
    delay = delay.long()
    span = int(delay.max().item()) + 1
    grp_key = (cls.long() * n_blocks + receiver.long() // n_r) * n_rows + (
        sender.long() * n_dirs + out_dir.long())
    order = torch.argsort(grp_key * span + delay)
    gk, d = grp_key[order], delay[order]
    first = torch.ones(gk.numel(), dtype=torch.bool, device=dev)
    first[1:] = gk[1:] != gk[:-1]
    group = torch.cumsum(first.long(), 0) - 1
    n_groups = int(group[-1].item()) + 1
    g_start = torch.nonzero(first).reshape(-1)
    pos = torch.arange(gk.numel(), device=dev) - g_start[group]
    assert int(pos.max().item()) < n_r, "more pairs than receiver slots in a group"
    widths = [w for w in WINDOW_CHOICES if w <= max_w] if width is None else [int(width)]
    covers = {w: _window_cover(d, group, pos, n_groups, n_r, w, align) for w in widths}
    n_wide = int(covers[widths[-1]][0].sum().item())
    width = next(w for w in widths if int(covers[w][0].sum().item()) <= 1.02 * n_wide)
    start, ebase = covers[width]
    rec = torch.cumsum(start.long(), 0) - 1
    n_rec = int(rec[-1].item()) + 1
    slot = (receiver.long() % n_r)[order]
    w = torch.zeros((n_rec, n_r), dtype=tdt, device=dev)
    w[rec, slot] = ff[order].to(tdt)
    rel = torch.full((n_rec, n_r), 255, dtype=torch.uint8, device=dev)
    rel[rec, slot] = (d - ebase).to(torch.uint8)
    head = torch.nonzero(start).reshape(-1)                 # first entry of each record
    meta = torch.stack([gk[head] % n_rows, ebase[head]], dim=1).to(torch.int32)
    recs = torch.cat([w.view(torch.uint8), rel, meta.contiguous().view(torch.uint8)],
                     dim=1).contiguous()
    assert recs.shape[1] == rec_bytes, (recs.shape, rec_bytes)
    counts = torch.bincount(gk[head] // n_rows, minlength=n_tiles)
    ent_ptr = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
    ent_ptr[1:] = torch.cumsum(counts, 0)
    return ent_ptr.contiguous(), recs, width


class EnergyHistogram:
    """(patch, direction, band, time) histogram in the padded-row device layout:
    ``data[(band * n_alloc + p) * D + dir, PAD + t]`` where ``p`` is the internal
    patch index of the pair tables (``tables.rank`` maps the caller's ids to it)."""

    def __init__(self, data, n_patches, n_dirs, n_bands, n_samples, pad, n_alloc=None,
                 tables=None, n_sources=None):
        self.data = data                    # (B * n_alloc * D, LD)
        self.n_patches, self.n_dirs = n_patches, n_dirs      # internal patch count
        self.n_bands, self.n_samples, self.pad = n_bands, n_samples, pad
        self.n_alloc = n_patches if n_alloc is None else n_alloc
        self.tables = tables
        self.n_sources = n_sources          # None: single source; else bands = S * B

    @property
    def ld(self):
        return self.data.shape[1]

    @property
    def rank(self):
        return None if self.tables is None else self.tables.rank

    def dense(self):
        """(N, D, B, T) in the reference's axis order and the caller's patch numbering
        (on the device; a strided view when no patch relabelling is active)."""
        v = self.data.view(self.n_bands, self.n_alloc, self.n_dirs, self.ld)
        v = v[:, :self.n_patches, :, self.pad:self.pad + self.n_samples]
        if self.rank is not None:
            v = v.index_select(1, self.rank)
        v = v.permute(1, 2, 0, 3)
        if self.n_sources is not None:      # (S, N, D, B, T)
            n, d, sb, t = v.shape
            v = v.reshape(n, d, self.n_sources, sb // self.n_sources, t).permute(2, 0, 1, 3, 4)
        return v

    def to_internal(self, per_patch, dim):
        return per_patch if self.tables is None else self.tables.to_internal(per_patch, dim)


class ExchangeWorkspace:
    """Device buffers of one exchange run: E_total, ping-pong E_a/E_b and G.  Passing the
    same workspace to several :func:`energy_exchange` calls re-uses the buffers (the
    histogram a call returns is then overwritten by the next one)."""

    def __init__(self, tables, n_samples, device, need_orders=True):
        t = tables
        self.t_pad, self.pad = _lib.exchange_layout(n_samples, t.max_delay, t.dtype)
        self.ld = self.t_pad + self.pad
        self.n_samples = n_samples
        self.sx = None
        if t.win_recs is not None and gather_kind() != "csr":
            # per-order driver of the tensor-memory gather: it owns the buffers
            from .distributed import ShardedExchange
            self.sx = ShardedExchange(t, n_samples, device, need_orders=need_orders, local=True)
            self.e_total, self.e_a, self.e_b, self.g = (
                self.sx.e_total, self.sx.e_a, self.sx.e_b, self.sx.g)
            return
        tdt = _lib.torch_dtype(t.dtype)
        rows = t.n_bands * t.n_patches * t.n_dirs
        self.e_total = torch.empty((rows, self.ld), dtype=tdt, device=device)
        if need_orders:
            self.e_a = torch.empty((rows, self.ld), dtype=tdt, device=device)
            self.e_b = torch.empty((rows, self.ld), dtype=tdt, device=device)
            # rows of empty segments are never written nor read
            self.g = torch.empty((t.n_bands * t.n_classes * t.n_patches, self.ld),
                                 dtype=tdt, device=device)
        else:
            self.e_a = self.e_b = self.g = None


def energy_exchange(tables, e0, delay0, n_samples, max_order, workspace=None):
    """``_energy_exchange`` (RadiosityFast.py:1073-1145) on the current device.

    e0: (N, D, B) device tensor; delay0: (N,) int32 source->patch delay bins.
    Returns an :class:`EnergyHistogram` holding sum_{k<=K} E_k.
    """
    if e0.dim() == 4:
        return _energy_exchange_batch(tables, e0, delay0, n_samples, max_order)
    if tables.win_recs is not None and gather_kind() != "csr":
        return _energy_exchange_orders(tables, e0, delay0, n_samples, max_order, workspace)
    t = tables
    tdt = _lib.torch_dtype(t.dtype)
    device = e0.device
    ws = workspace or ExchangeWorkspace(t, n_samples, device, need_orders=max_order >= 1)
    e0, delay0 = t.to_internal(e0), t.to_internal(delay0)
    e0 = e0.to(tdt).contiguous()
    delay0 = delay0.to(torch.int32).contiguous()
    use_tiles = t.recs is not None and gather_kind() != "csr"
    _lib.call("spb_energy_exchange", e0, delay0, t.seg_ptr, t.src, t.wgt, t.dly,
              t.ent_ptr if use_tiles else None, t.recs if use_tiles else None, t.coef,
              t.n_patches, t.n_classes, t.n_dirs, t.n_bands, n_samples, ws.t_pad, ws.pad,
              int(max_order), ws.e_total, ws.e_a, ws.e_b, ws.g, _lib.I32(t.dtype),
              _lib.stream_ptr())
    return EnergyHistogram(ws.e_total, t.n_patches, t.n_dirs, t.n_bands, n_samples, ws.pad,
                           tables=t)


def _energy_exchange_orders(tables, e0, delay0, n_samples, max_order, workspace=None):
    """One source through the per-order driver (gather + mix launched from Python) --
    the path of the tensor-memory gather (``spb_energy_exchange`` drives the tiled and CSR
    kernels).  Single device, whatever process group may be initialised."""
    from .distributed import ShardedExchange
    if workspace is not None and workspace.sx is not None and (
            workspace.sx.e_a is not None or max_order < 1):
        sx = workspace.sx
    else:
        sx = ShardedExchange(tables, n_samples, e0.device, need_orders=max_order >= 1,
                             local=True)
    sx.init(e0, delay0)
    return sx.run(max(0, int(max_order)))


def _energy_exchange_batch(tables, e0, delay0, n_samples, max_order):
    """Several sources at once: e0 (S, N, D, B), delay0 (S, N).  The sources ride along
    as extra bands (PairTables.tiled), so every kernel launch covers all of them.
    Returns an EnergyHistogram whose dense() is (S, N, D, B, T)."""
    from .distributed import ShardedExchange
    n_src = e0.shape[0]
    sx = ShardedExchange(tables.tiled(n_src), n_samples, e0.device,
                         need_orders=max_order >= 1, local=True)
    sx.init(e0, delay0)
    hist = sx.run(max(0, int(max_order)))
    hist.n_sources = n_src
    return hist


STAGED_RECEIVERS, STAGED_BINS = 8, 4     # per CTA / per thread (csrc/exchange.cu)


def collect_kind(hist, n_rcv):
    """Which collection kernel sums the patches: ``("staged", stages)`` --
    `k_collect_staged`, every histogram row staged once in shared memory for a group of 8
    receivers (diffuse scenes with many receivers, BASELINE config 3) -- or ``("direct",)``,
    `k_collect_partial` (one row read per receiver).  ``SPB_COLLECT=direct`` or
    ``staged[:stages]`` forces one (cross-check tests, tools/sweep_collect.py)."""
    env = os.environ.get("SPB_COLLECT", "").split(":")
    esize = hist.data.element_size()
    stages = int(env[1]) if env[0] == "staged" and len(env) > 1 else 0
    vec = 16 // esize
    chunk = 256 * STAGED_BINS
    row2 = -(-hist.n_samples // vec) * vec + -(-hist.n_samples // chunk) * chunk
    stage_bytes = (row2 + STAGED_RECEIVERS) * esize + 4 * STAGED_RECEIVERS
    fits = hist.n_dirs == 1 and max(2, stages) * stage_bytes <= 220 * 1024
    if env[0] == "staged":
        if not fits:
            raise _lib.SparrowB200Error("SPB_COLLECT=staged needs one direction per patch and "
                                        "histogram rows that fit shared memory")
        return ("staged", stages)
    if env[0] == "direct" or not fits or n_rcv < 4:
        return ("direct",)
    return ("staged", stages)


def collect_mono(hist, rdir, shift, scale, n_split=None):
    """Sum of all patch histograms at each receiver, (R, B, T)
    (``collect_energy_receiver_mono``, RadiosityFast.py:570-602 / :1148-1185).

    rdir, shift: (R, N) int32; scale: (R, N, B).
    """
    n_rcv = rdir.shape[0]
    tdt = hist.data.dtype
    code = _lib.dtype_code(tdt)
    rdir, shift, scale = (hist.to_internal(x, 1) for x in (rdir, shift, scale))
    if hist.n_sources is not None:          # the same receiver factors for every source
        scale = scale.repeat(1, 1, hist.n_sources)
    kind = collect_kind(hist, n_rcv)
    if n_split is None:
        n_split = max(1, min(64, hist.n_patches // 256))
        if kind[0] == "staged":             # ~8 waves of CTAs (2 per SM), >= 64 patches each
            ctas = (hist.n_bands * -(-n_rcv // STAGED_RECEIVERS)
                    * -(-hist.n_samples // (256 * STAGED_BINS)))
            n_split = max(1, min(hist.n_patches // 64, -(-8 * 2 * 148 // ctas)))
    out = torch.empty((n_rcv, hist.n_bands, hist.n_samples), dtype=tdt,
                      device=hist.data.device)
    # grid.y carries receiver*band: batch the receivers if there are many
    step = max(1, 65535 // hist.n_bands)
    for r0 in range(0, n_rcv, step):
        r1 = min(n_rcv, r0 + step)
        partial = torch.empty((n_split, r1 - r0, hist.n_bands, hist.n_samples), dtype=tdt,
                              device=hist.data.device)
        if kind[0] == "staged":
            _lib.call("spb_collect_mono_staged", hist.data, shift[r0:r1].contiguous(),
                      scale[r0:r1].to(tdt).contiguous(), r1 - r0, hist.n_patches, hist.n_alloc,
                      hist.n_bands, hist.n_samples, hist.ld, hist.pad, out[r0:r1], partial,
                      n_split, _lib.I32(kind[1]), _lib.I32(code), _lib.stream_ptr())
            continue
        _lib.call("spb_collect_mono", hist.data, rdir[r0:r1].contiguous(),
                  shift[r0:r1].contiguous(), scale[r0:r1].to(tdt).contiguous(), r1 - r0,
                  hist.n_patches, hist.n_alloc, hist.n_dirs, hist.n_bands, hist.n_samples,
                  hist.ld,
                  hist.pad, out[r0:r1], partial, n_split, _lib.I32(code),
                  _lib.stream_ptr())
    if hist.n_sources is not None:          # (S, R, B, T)
        out = out.reshape(n_rcv, hist.n_sources, -1, hist.n_samples).permute(1, 0, 2, 3)
    return out


def collect_patchwise(hist, rdir, shift, scale):
    """Per-patch histograms at each receiver, (R, N, B, T) -- (S, R, N, B, T) for a batch
    of sources (``collect_energy_receiver_patchwise``, RadiosityFast.py:660-752)."""
    n_rcv = rdir.shape[0]
    tdt = hist.data.dtype
    code = _lib.dtype_code(tdt)
    rdir, shift, scale = (hist.to_internal(x, 1) for x in (rdir, shift, scale))
    if hist.n_sources is not None:          # the same receiver factors for every source
        scale = scale.repeat(1, 1, hist.n_sources)
    out = torch.empty((n_rcv, hist.n_patches, hist.n_bands, hist.n_samples), dtype=tdt,
                      device=hist.data.device)
    _lib.call("spb_collect_patchwise", hist.data, rdir.contiguous(), shift.contiguous(),
              scale.to(tdt).contiguous(), n_rcv, hist.n_patches, hist.n_alloc, hist.n_dirs,
              hist.n_bands, hist.n_samples, hist.ld, hist.pad, out, _lib.I32(code),
              _lib.stream_ptr())
    if hist.rank is not None:
        out = out.index_select(1, hist.rank)
    if hist.n_sources is not None:          # (S, R, N, B, T)
        r, n = out.shape[:2]
        out = out.reshape(r, n, hist.n_sources, -1, hist.n_samples).permute(2, 0, 1, 3, 4)
    return out


def energy_exchange_host(tables, e0_host, distance0_host, speed_of_sound, dt, n_samples,
                         max_order, out_host=None, workspace=None):
    """``_energy_exchange`` (RadiosityFast.py:1073-1145) with HOST buffers in and out:
    the call a maintainer binds in place of the numba kernel.

    e0_host: (N, D, B) float64, distance0_host: (N,) float64 source->patch distances
    (torch CPU tensors, ideally pinned, or numpy arrays).  The full (N, D, B, T)
    histogram is written to ``out_host`` (allocated pinned when None) and returned.
    """
    from . import bake
    dev = tables.seg_ptr.device
    e0 = torch.as_tensor(e0_host).to(dev, non_blocking=True)
    d0 = torch.as_tensor(distance0_host).to(dev, non_blocking=True)
    delay0 = bake.delay_bins(d0, speed_of_sound, dt)
    hist = energy_exchange(tables, e0, delay0, n_samples, max_order, workspace=workspace)
    dense = hist.dense()
    if out_host is None:
        out_host = torch.empty(dense.shape, dtype=dense.dtype).pin_memory()
    out_host.copy_(dense)
    return out_host
