"""Host-side index bookkeeping of the exchange on CPU tensors: CSR segments and the
tile records of the TMA gather are lossless re-encodings of the directed pair list;
patch renumbering and source tiling keep it intact.  (No kernels run here; the
library is only asked for its tile geometry.)"""
import numpy as np
import pytest
import torch

from sparrowpy_b200 import _lib, exchange, geometry, scenes


def random_pairs(seed, n=61, d=3, c=4, m=2500, t_len=200):
    gen = torch.Generator().manual_seed(seed)
    sender = torch.randint(0, n, (m,), generator=gen)
    receiver = torch.randint(0, n, (m,), generator=gen)
    _, first = np.unique((sender * n + receiver).numpy(), return_index=True)
    sel = torch.from_numpy(np.sort(first))
    sender, receiver = sender[sel], receiver[sel]
    m = sender.numel()
    ff = torch.rand(m, generator=gen, dtype=torch.float64) + 0.01
    delay = torch.randint(0, 260, (m,), generator=gen)      # some >= t_len: dropped
    out_dir = torch.randint(0, d, (m,), generator=gen)
    cls = torch.randint(0, c, (m,), generator=gen)
    coef = torch.rand((c, d, 2), generator=gen, dtype=torch.float64)
    return sender, receiver, ff, delay, out_dir, cls, coef, n, t_len


def decode_records(t, dtype=np.float64):
    """(class, receiver, src_row, delay, weight) tuples encoded in the tile records."""
    n_r, bucket, rec_bytes = exchange.tile_geometry(t.dtype)
    wsize = 8 if t.dtype == _lib.F64 else 4
    recs = t.recs.numpy()
    ent = t.ent_ptr.numpy()
    n_blocks = -(-t.n_patches // n_r)
    out = []
    for tile in range(len(ent) - 1):
        c, jb = divmod(tile, n_blocks)
        for e in range(ent[tile], ent[tile + 1]):
            raw = recs[e]
            w = raw[:n_r * wsize].view(np.float64 if wsize == 8 else np.float32)
            rel = raw[n_r * wsize:n_r * wsize + n_r]
            src, packed = raw[n_r * wsize + n_r:].view(np.int32)
            dmin, mask = int(packed) & 0xffffff, (int(packed) >> 24) & 0xff
            assert dmin % bucket == 0
            cur = None
            for s in range(n_r):
                if w[s] != 0:
                    out.append((c, jb * n_r + s, int(src), dmin + int(rel[s]), float(w[s])))
                    # the reload mask marks exactly the changes of shift among used slots
                    assert bool(mask >> s & 1) == (cur != rel[s])
                    cur = rel[s]
                else:
                    # empty slots never trigger a reload and repeat the current shift
                    assert not (mask >> s & 1)
                    assert rel[s] == (0 if cur is None else cur)
    return sorted(out)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_csr_and_tile_records_encode_the_same_pairs(dtype):
    sender, receiver, ff, delay, out_dir, cls, coef, n, t_len = random_pairs(1)
    t = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n, t_len,
                                   dtype, gather="tma")
    keep = delay < t_len
    assert t.n_directed == sender.numel() and t.src.numel() == int(keep.sum())
    assert t.max_delay == int(delay[keep].max())
    cast = np.float64 if dtype == "f64" else np.float32
    want = sorted((int(c), int(r), int(s) * t.n_dirs + int(o), int(dl), float(cast(w)))
                  for c, r, s, o, dl, w in zip(cls[keep], receiver[keep], sender[keep],
                                               out_dir[keep], delay[keep], ff[keep]))
    # CSR
    seg_ptr = t.seg_ptr.numpy()
    got = []
    for seg in range(len(seg_ptr) - 1):
        c, j = divmod(seg, n)
        for q in range(seg_ptr[seg], seg_ptr[seg + 1]):
            got.append((c, j, int(t.src[q]), int(t.dly[q]), float(t.wgt[q])))
    assert sorted(got) == want
    # tile records
    assert decode_records(t) == want


def test_renumbering_and_source_tiling():
    sender, receiver, ff, delay, out_dir, cls, coef, n, t_len = random_pairs(2)
    gen = np.random.default_rng(0)
    n_int = n + 11
    rank = torch.from_numpy(gen.permutation(n_int)[:n].astype(np.int64))
    t = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n, t_len,
                                   "f64", rank=rank, n_internal=n_int, gather="tma")
    assert (t.n_patches, t.n_user) == (n_int, n)
    keep = delay < t_len
    want = sorted((int(c), int(rank[r]), int(rank[s]) * t.n_dirs + int(o), int(dl), float(w))
                  for c, r, s, o, dl, w in zip(cls[keep], receiver[keep], sender[keep],
                                               out_dir[keep], delay[keep], ff[keep]))
    assert decode_records(t) == want
    x = torch.arange(n, dtype=torch.float64)
    xi = t.to_internal(x)
    assert xi.shape == (n_int,) and torch.equal(xi[rank], x) and float(xi.sum()) == float(x.sum())
    t3 = t.tiled(3)
    assert t3.n_bands == 3 * t.n_bands and t3.coef.shape == (t.n_classes, t.n_dirs, 6)
    assert torch.equal(t3.coef[:, :, 2:4], t.coef) and t3.recs is t.recs


def test_compact_order_is_a_balanced_numbering():
    walls = scenes.street_canyon(0, 0.25)
    pts, ids = geometry.process_patches(np.array([w[0] for w in walls]), 1.0)
    for shards in (1, 2, 4, 8):
        rank, n_int = geometry.compact_patch_order(pts, ids, n_shards=shards)
        assert len(np.unique(rank)) == len(ids) and rank.max() < n_int
        assert n_int % (8 * shards) == 0
        size = n_int // shards
        per_shard = np.bincount(rank // size, minlength=shards)
        assert per_shard.max() - per_shard.min() <= max(16, 0.2 * per_shard.mean())
    # tiles of 8 consecutive internal indices are spatially compact
    rank, n_int = geometry.compact_patch_order(pts, ids)
    cen = pts.mean(axis=1)
    inv = -np.ones(n_int, int)
    inv[rank] = np.arange(len(rank))
    ext = []
    for k in range(0, n_int, 8):
        m = inv[k:k + 8]
        m = m[m >= 0]
        if len(m) > 1:
            ext.append(np.ptp(cen[m], axis=0).max())
    assert np.mean(ext) <= 3.0 + 1e-9        # 2 x 4 patches of 1 m: extent <= 3 m


def decode_window_records(ent_ptr, recs, width, n_patches, n_r=8):
    """(class, receiver, src_row, delay, weight) tuples encoded in the window records."""
    recs, ent = recs.numpy(), ent_ptr.numpy()
    n_blocks = -(-n_patches // n_r)
    out = []
    for tile in range(len(ent) - 1):
        c, jb = divmod(tile, n_blocks)
        prev = None
        for e in range(ent[tile], ent[tile + 1]):
            raw = recs[e]
            w = raw[:64].view(np.float64)
            rel = raw[64:72]
            src, dbase = (int(x) for x in raw[72:].view(np.int32))
            assert dbase % 2 == 0 and dbase >= 0
            assert (src, dbase) > prev if prev is not None else True     # sorted, no duplicates
            prev = (src, dbase)
            used = rel != 255
            assert used.any() and (rel[used] <= width).all() and (w[~used] == 0).all()
            for s in np.nonzero(used)[0]:
                out.append((c, jb * n_r + int(s), src, dbase + int(rel[s]), float(w[s])))
    return sorted(out)


@pytest.mark.parametrize("align", [2, 4])
@pytest.mark.parametrize("width", [None, 4, 10])
def test_window_records_encode_the_same_pairs(width, align):
    sender, receiver, ff, delay, out_dir, cls, coef, n, t_len = random_pairs(3)
    keep = delay < t_len
    args = [x[keep] for x in (sender, receiver, ff, delay, out_dir, cls)]
    ent_ptr, recs, w = exchange.build_window_records(*args, n, 3, 4, _lib.F64, width=width,
                                                     align=align)
    assert w in exchange.WINDOW_CHOICES and (width is None or w == width)
    dbase = recs.numpy()[:, 76:80].copy().view(np.int32).reshape(-1)
    first = np.array([r[64:72][r[64:72] != 255].min() for r in recs.numpy()])
    # sector-aligned rows wherever a non-negative base with that residue exists
    ok = (dbase + w) % align == 0
    assert ok[dbase + first >= align].all() and (dbase % 2 == 0).all()
    want = sorted((int(c), int(r), int(s) * 3 + int(o), int(dl), float(x))
                  for s, r, x, dl, o, c in zip(*args))
    assert decode_window_records(ent_ptr, recs, w, n) == want


def test_window_records_merge_neighbouring_delays():
    """8 receivers of one tile seeing the same sender row with delays inside one window
    become a single record; a delay outside it starts a second one."""
    n, d = 16, 1
    receiver = torch.arange(8)
    sender = torch.full((8,), 12)
    ff = torch.arange(1, 9, dtype=torch.float64)
    delay = torch.tensor([41, 40, 43, 47, 50, 45, 44, 61])
    zeros = torch.zeros(8, dtype=torch.int64)
    ent_ptr, recs, w = exchange.build_window_records(sender, receiver, ff, delay, zeros, zeros,
                                                     n, d, 1, _lib.F64, width=10)
    assert ent_ptr.tolist() == [0, 2, 2] and recs.shape == (2, 80)
    raw = recs.numpy()
    assert raw[0, 64:72].tolist() == [1, 0, 3, 7, 10, 5, 4, 255]
    assert raw[0, 72:].view(np.int32).tolist() == [12, 40]
    assert raw[1, 64:72].tolist() == [255] * 7 + [1] and raw[1, 72:].view(np.int32).tolist() == [12, 60]
    assert raw[1, :64].view(np.float64).tolist() == [0.0] * 7 + [8.0]


def test_window_swizzle_is_conflict_free():
    """The shared-memory layout of k_gather_win: a lane reads 16-byte chunks
    v = 128*warp + 4*lane + q (q = window chunk); with chunk bits 0-1 ^= bits 3-4 the 8
    lanes of every quarter warp (one LDS.128 wavefront) fall into 8 different 16-byte
    bank groups for every q, and the map is a permutation inside each 512-byte block."""
    swz = lambda v: v ^ ((v >> 3) & 3)    # noqa: E731
    assert sorted(swz(v) for v in range(64)) == list(range(64))
    for warp in range(8):
        for q in range(9):
            for quarter in range(4):
                groups = {swz(128 * warp + 4 * lane + q) % 8
                          for lane in range(8 * quarter, 8 * quarter + 8)}
                assert len(groups) == 8
    # variant with 4 bins per lane: lane stride 2 chunks, bit 0 ^= bit 3
    swz4 = lambda v: v ^ ((v >> 3) & 1)   # noqa: E731
    assert sorted(swz4(v) for v in range(64)) == list(range(64))
    for warp in range(16):
        for q in range(7):
            for quarter in range(4):
                groups = {swz4(64 * warp + 2 * lane + q) % 8
                          for lane in range(8 * quarter, 8 * quarter + 8)}
                assert len(groups) == 8


def test_visible_pairs_scans_the_matrix_in_row_blocks():
    """bake.visible_pairs must not depend on the block size (torch.nonzero is limited to
    2^31 elements per call; config 5 has a 10^10-element visibility matrix)."""
    from sparrowpy_b200 import bake
    gen = torch.Generator().manual_seed(0)
    vis = torch.triu(torch.rand((57, 57), generator=gen) < 0.2, 1)
    whole = bake.visible_pairs(vis)
    assert whole.dtype == torch.int32 and whole.shape[1] == 2
    assert torch.equal(whole.long(), torch.nonzero(vis))
    for block in (1, 57, 100, 57 * 5 + 3, 57 * 57):
        assert torch.equal(bake.visible_pairs(vis, max_block_elems=block), whole)


def decode_window_records(t):
    """(class, receiver, src_row, delay, weight) tuples encoded in the window records as the
    kernel reads them (csrc/exchange_tmem.cu, exchange.device_window_records):
    {w[8] f64, 2 * rel[8] u8, src i32, dbase i32}; an empty slot has w = 0 and offset 0."""
    recs, ent = t.win_recs.numpy(), t.win_ptr.numpy()
    n_blocks = -(-t.n_patches // 8)
    out, n_null = [], 0
    for tile in range(len(ent) - 1):
        c, jb = divmod(tile, n_blocks)
        for e in range(ent[tile], ent[tile + 1]):
            raw = recs[e]
            w, off = raw[:64].view(np.float64), raw[64:72]
            src, dbase = raw[72:].view(np.int32)
            assert dbase % 2 == 0 and dbase >= 0
            if np.all(w == 0):
                assert np.all(off == 0) and src == 0 and dbase == 0      # padding record
                n_null += 1
            for s in range(8):
                if w[s] != 0:
                    assert off[s] % 2 == 0 and off[s] // 2 <= t.win_w
                    out.append((c, jb * 8 + s, int(src), int(dbase) + int(off[s]) // 2,
                                float(w[s])))
                else:
                    assert off[s] == 0
    return sorted(out), n_null


def test_window_records_encode_the_same_pairs(gather="tmem"):
    """The window records are a lossless encoding of the pair list; the tensor-memory
    kernel's lists are padded to its batch size with null records."""
    sender, receiver, ff, delay, out_dir, cls, coef, n, t_len = random_pairs(3)
    t = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n, t_len,
                                   "f64", gather=gather)
    assert t.recs is None and t.win_recs is not None and t.win_w in exchange.WINDOW_CHOICES
    keep = delay < t_len
    want = sorted((int(c), int(r), int(s) * t.n_dirs + int(o), int(dl), float(w))
                  for c, r, s, o, dl, w in zip(cls[keep], receiver[keep], sender[keep],
                                               out_dir[keep], delay[keep], ff[keep]))
    got, n_null = decode_window_records(t)
    assert got == want
    counts = (t.win_ptr[1:] - t.win_ptr[:-1]).numpy()
    batch = exchange.tmem_batch()
    assert np.all(counts % batch == 0) and n_null < batch * len(counts)


def test_csr_lists_of_index_ranges():
    """bake._csr_lists: element k is a member of the lists first[k]..last[k], order kept."""
    from sparrowpy_b200 import bake
    first, last = np.array([0, 2, 1, 3, 2]), np.array([1, 2, 3, 2, 2])     # element 3: empty range
    ptr, items = bake._csr_lists(first, last, 5, np.array([10, 11, 12, 13, 14], np.int32))
    lists = [items[ptr[i]:ptr[i + 1]].tolist() for i in range(5)]
    assert lists == [[10], [10, 12], [11, 12, 14], [12], []]


def test_collection_kernel_choice_is_host_logic(monkeypatch):
    """exchange.collect_kind needs no device: staged for diffuse histograms with >= 4 receivers
    whose doubled rows fit shared memory, direct otherwise; SPB_COLLECT overrides."""
    import torch
    from sparrowpy_b200 import exchange
    monkeypatch.delenv("SPB_COLLECT", raising=False)
    hist = exchange.EnergyHistogram(torch.zeros((4, 1056), dtype=torch.float64), 4, 1, 1, 1000, 32)
    assert exchange.collect_kind(hist, 64) == ("staged", 0)
    assert exchange.collect_kind(hist, 3) == ("direct",)
    hist.n_samples = 13000                                   # 2 x (2 x 13000 + ...) x 8 B > 220 KB
    assert exchange.collect_kind(hist, 64) == ("direct",)
    hist.n_samples, hist.n_dirs = 1000, 16
    assert exchange.collect_kind(hist, 64) == ("direct",)
    hist.n_dirs = 1
    monkeypatch.setenv("SPB_COLLECT", "direct")
    assert exchange.collect_kind(hist, 64) == ("direct",)
    monkeypatch.setenv("SPB_COLLECT", "staged:4")
    assert exchange.collect_kind(hist, 1) == ("staged", 4)


@pytest.mark.parametrize("esize", [8, 4])
def test_doubled_row_addressing_of_the_staged_collection(esize):
    """Model of k_collect_staged's shared-memory row (csrc/exchange.cu): copy A at
    [t_al - T, t_al), copy B at [t_al, t_al + T), read at t_al - shift + t.  For every bin
    t < T and shift in [0, T) that is E[(t - shift) mod T] -- the reference's np.roll
    (RadiosityFast.py:1181-1183) -- and every thread of the CTA, tail threads included, stays
    inside the stage."""
    rng = np.random.default_rng(esize)
    vec = 16 // esize
    for n_samples in [1, 2, 3, 7, 64, 255, 256, 333, 1000, 1024, 1025, 2050]:
        t_al = -(-n_samples // vec) * vec
        n_chunks = -(-n_samples // 1024)
        row_elems = n_chunks * 1024
        row2 = t_al + row_elems
        e = rng.uniform(1, 2, n_samples)
        stage = np.full(row2, np.nan)
        stage[2 * t_al:] = 0.0                         # zeroed once per stage
        stage[t_al - n_samples:t_al] = e               # copy A
        n_copy = min(t_al, row2 - t_al)                # 16-byte pieces of copy B (may spill)
        stage[t_al:t_al + n_samples] = e
        stage[t_al + n_samples:t_al + n_copy] = 7.0    # row padding carried along by the pieces
        t_all = np.arange(row_elems)                   # chunk * 1024 + tid + 256 q
        for shift in {0, 1, n_samples // 2, n_samples - 1}:
            idx = t_al - shift + t_all
            assert idx.min() >= 0 and idx.max() < row2
            got = stage[idx[:n_samples]]
            assert np.array_equal(got, np.roll(e, shift))
            assert not np.isnan(stage[idx]).any()      # tail threads read initialised memory
