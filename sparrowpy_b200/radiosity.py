"""``DirectionalRadiosityFast`` -- drop-in for the reference class of the same name
(reference sparrowpy/classes/RadiosityFast.py:15-968) with every numba kernel
replaced by an sm_100a CUDA kernel behind the C ABI of include/sparrow_b200.h.

Same method names, keyword arguments, error behaviour and result types.  State
lives on the GPU; the numpy attributes the reference exposes
(``_visibility_matrix``, ``_form_factors``, ``_form_factors_tilde``,
``_energy_exchange_etc`` ...) are materialised on access.  ``form_factors_tilde``
is kept *factored* (per-pair form factor x small BRDF table) because the dense
(N, N, D, B) tensor does not scale; the dense tensor is only
built when the attribute is read.

There is no CPU fallback: without the CUDA library and a CUDA device every compute
method raises ``SparrowB200Error``.
"""
import os

import numpy as np
import torch

from . import _lib, bake, exchange, geometry, pyfar_shim, sound_object

try:  # accept real pyfar objects when the package is present
    import pyfar as _pf
    _COORD_TYPES = (pyfar_shim.Coordinates, _pf.Coordinates)
    _TimeData, _Coordinates, _FrequencyData = (
        _pf.TimeData, _pf.Coordinates, _pf.FrequencyData)
except Exception:  # noqa: BLE001  (pyfar missing or broken)
    _pf = None
    _COORD_TYPES = (pyfar_shim.Coordinates,)
    _TimeData, _Coordinates, _FrequencyData = (
        pyfar_shim.TimeData, pyfar_shim.Coordinates, pyfar_shim.FrequencyData)


def _np(t):
    return t.detach().cpu().numpy()


class DirectionalRadiosityFast:
    """Radiosity object for directional scattering coefficients (B200 path)."""

    def __init__(
            self, walls_points, walls_normal, walls_up_vector, patches_points,
            n_patches, patch_to_wall_ids, visibility_matrix=None,
            visible_patches=None, form_factors=None, form_factors_tilde=None,
            frequencies=None, brdf=None, brdf_index=None,
            brdf_incoming_directions=None, brdf_outgoing_directions=None,
            patch_2_brdf_outgoing_index=None, air_attenuation=None,
            speed_of_sound=None, etc_time_resolution=None, etc_duration=None,
            distance_patches_to_source=None, energy_init_source=None,
            energy_exchange_etc=None, device=None, dtype="f64"):
        _lib.dtype_code(dtype)
        self._dtype = dtype
        self._device = torch.device(device if device is not None else "cuda")
        self._walls_points = np.atleast_3d(np.asarray(walls_points, float))
        self._walls_up_vector = np.atleast_2d(np.asarray(walls_up_vector, float))
        self._walls_normal = np.atleast_2d(np.asarray(walls_normal, float))
        self._patches_points = np.atleast_3d(np.asarray(patches_points, float))
        self._n_patches = int(n_patches)
        self._patch_to_wall_ids = np.atleast_1d(np.array(patch_to_wall_ids, dtype=int))

        self._frequencies = None if frequencies is None else np.array(frequencies)
        self._brdf = None if brdf is None else [np.array(b) for b in brdf]
        self._brdf_index = (None if brdf_index is None
                            else np.atleast_1d(np.array(brdf_index, dtype=np.int64)))
        self._brdf_incoming_directions = _object_array(brdf_incoming_directions)
        self._brdf_outgoing_directions = _object_array(brdf_outgoing_directions)
        self._air_attenuation = (None if air_attenuation is None
                                 else np.array(air_attenuation))
        self._speed_of_sound = None if speed_of_sound is None else float(speed_of_sound)
        self._etc_time_resolution = (None if etc_time_resolution is None
                                     else float(etc_time_resolution))
        self._etc_duration = None if etc_duration is None else float(etc_duration)
        self._source = None

        # device state -------------------------------------------------------
        self._d = {}                 # geometry tensors (built lazily)
        self._baked = None           # dict of baked pair tensors
        self._tables = None          # (key, PairTables) of the last exchange
        self._hist = None            # EnergyHistogram (device)
        self._e0_dev = None
        self._d0_dev = None
        self._source_vis_dev = None
        # host copies handed in through the constructor (checkpoint resume)
        self._host = dict(
            visibility_matrix=None if visibility_matrix is None
            else np.array(visibility_matrix),
            visible_patches=None if visible_patches is None
            else np.array(visible_patches),
            form_factors=None if form_factors is None else np.array(form_factors),
            form_factors_tilde=None if form_factors_tilde is None
            else np.array(form_factors_tilde),
            patch_2_brdf_outgoing_index=None if patch_2_brdf_outgoing_index is None
            else np.array(patch_2_brdf_outgoing_index, dtype=np.int64),
            distance_patches_to_source=None if distance_patches_to_source is None
            else np.array(distance_patches_to_source),
            energy_init_source=None if energy_init_source is None
            else np.array(energy_init_source),
            energy_exchange_etc=None if energy_exchange_etc is None
            else np.array(energy_exchange_etc))
        self.check()

    # ------------------------------------------------------------------
    # validation: reference RadiosityFast.py:211-337
    # ------------------------------------------------------------------
    def check(self):
        """Check the input data for consistency."""
        n_walls = self._walls_points.shape[0]
        if self._walls_points.ndim != 3 or self._walls_points.shape[2] != 3:
            raise ValueError("Walls need to be of shape (n_walls, n_points, 3)")
        if self._walls_up_vector.shape != (n_walls, 3):
            raise ValueError("Up vector of walls need to be of shape (n_walls, 3)")
        if self._walls_normal.shape != (n_walls, 3):
            raise ValueError("Normal of walls need to be of shape (n_walls, 3)")
        if (self._patches_points.shape[0] != self.n_patches) or \
                (self._patches_points.shape[2] != 3):
            raise ValueError("Patches need to be of shape (n_patches, n_points, 3)")
        if self._patch_to_wall_ids.shape != (self.n_patches,):
            raise ValueError("patch_to_wall_ids need to be of shape (n_patches,)")
        ids = set(self._patch_to_wall_ids.tolist())
        if ids != set(range(n_walls)):
            raise ValueError(
                "patch_to_wall_ids does contain other ids than range(n_walls)")
        n_bins = 1
        if self._frequencies is not None:
            if len(self._frequencies.shape) != 1:
                raise ValueError("Frequencies need to be of shape (n_bins,)")
            n_bins = self._frequencies.size
        h = self._host
        if h["form_factors"] is not None:
            if h["form_factors"].shape != (self.n_patches, self.n_patches):
                raise ValueError(
                    "form_factors need to be of shape (n_patches, n_patches)")
        n_out = 1
        if self._brdf_index is not None and len(self._brdf_index) != n_walls:
            raise ValueError("brdf_index need to be of shape (n_walls,)")
        for name in ("_brdf_incoming_directions", "_brdf_outgoing_directions"):
            dirs = getattr(self, name)
            if dirs is not None and any(not isinstance(i, _COORD_TYPES) for i in dirs):
                raise ValueError(
                    f"{name[1:]} need to be a list of type pf.Coordinates")
        if self._brdf_outgoing_directions is not None:
            n_out = self._brdf_outgoing_directions[0].csize
        if h["form_factors_tilde"] is not None:
            if h["form_factors_tilde"].shape != (
                    self.n_patches, self.n_patches, n_out, n_bins):
                raise ValueError(
                    "form_factors_tilde need to be of shape "
                    "(n_patches, n_patches, n_outgoing_directions, n_bins)")
        if self._air_attenuation is not None:
            if len(self._air_attenuation.shape) != 1 or \
                    self._air_attenuation.shape[0] != n_bins:
                raise ValueError("Air attenuation need to be of shape (n_bins,)")
        if self._speed_of_sound is not None and self._speed_of_sound <= 0:
            raise ValueError("Speed of sound must be positive and non-zero")
        if self._etc_time_resolution is not None and self._etc_time_resolution <= 0:
            raise ValueError("Time resolution must be positive and non-zero")
        if self._etc_duration is not None and self._etc_duration <= 0:
            raise ValueError("Duration must be positive and non-zero")
        if h["distance_patches_to_source"] is not None:
            if h["distance_patches_to_source"].shape != (self.n_patches,):
                raise ValueError(
                    "distance_patches_to_source need to be of shape (n_patches,)")
        if h["energy_init_source"] is not None:
            if h["energy_init_source"].shape != (self.n_patches, n_out, n_bins):
                raise ValueError(
                    "energy_init_source need to be of shape "
                    "(n_patches, n_outgoing_directions, n_bins)")
        if h["energy_exchange_etc"] is not None:
            n_samples = int(self._etc_duration / self._etc_time_resolution)
            if h["energy_exchange_etc"].shape != (
                    self.n_patches, n_out, n_bins, n_samples):
                raise ValueError(
                    "energy_exchange_etc need to be of shape "
                    "(n_patches, n_outgoing_directions, n_bins, n_samples)")

    # ------------------------------------------------------------------
    @classmethod
    def from_polygon(cls, polygon_list, patch_size, device=None, dtype="f64"):
        """Create a radiosity object from wall polygons (RadiosityFast.py:340-367).

        ``device`` / ``dtype`` ('f64' reference precision, 'f32' histograms in
        single precision) are additions of this implementation.
        """
        walls_points = np.array([p.pts for p in polygon_list])
        walls_normal = np.array([p.normal for p in polygon_list])
        walls_up_vector = np.array([p.up_vector for p in polygon_list])
        patches_points, patch_to_wall_ids = geometry.process_patches(
            walls_points, patch_size)
        return cls(walls_points, walls_normal, walls_up_vector, patches_points,
                   patches_points.shape[0], patch_to_wall_ids, device=device,
                   dtype=dtype)

    # ------------------------------------------------------------------
    # device geometry
    # ------------------------------------------------------------------
    def _geom(self):
        if not self._d:
            _lib.require_cuda()
            dev = self._device
            t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
            self._d = dict(
                points=t(self._patches_points), center=t(self.patches_center),
                normal=t(self.patches_normal), area=t(self.patches_area),
                wall_ids=t(self._patch_to_wall_ids.astype(np.int64)),
                walls_points=t(self._walls_points), walls_normal=t(self._walls_normal))
        return self._d

    def _brdf_tables(self):
        """(vi (W,S,3), vo (W,D,3), brdf (n_brdf,S,D,B), brdf_index (W,)) as numpy."""
        vi = np.array([s.cartesian.reshape(-1, 3) for s in self._brdf_incoming_directions])
        vo = np.array([s.cartesian.reshape(-1, 3) for s in self._brdf_outgoing_directions])
        n_bins = self.n_bins
        brdf = np.array([np.real(np.asarray(b)).reshape(vi.shape[1], vo.shape[1], n_bins)
                         for b in self._brdf])
        return vi, vo, brdf, np.asarray(self._brdf_index, np.int64)

    # ------------------------------------------------------------------
    # bake: RadiosityFast.py:369-433
    # ------------------------------------------------------------------
    def bake_geometry(self):
        """Bake the geometry: visibility, form factors, BRDF direction tables."""
        g = self._geom()
        # blockers = the patches themselves (RadiosityFast.py:374-375); grouped by wall
        # the conjunction over blockers is evaluated hierarchically (same result)
        if os.environ.get("SPB_VISIBILITY", "grouped") == "brute":
            vis = bake.visibility_p2p(g["center"], g["normal"], g["points"])
        else:
            vis = bake.visibility_p2p_grouped(g["center"], g["normal"], g["points"],
                                              self._patch_to_wall_ids)
        pairs = bake.visible_pairs(vis)
        ff, _ = bake.form_factors(g["points"], g["normal"], g["area"], pairs)
        self._bake_pairs(vis, pairs, ff)
        self._tables = None
        for k in ("visibility_matrix", "visible_patches", "form_factors",
                  "form_factors_tilde", "patch_2_brdf_outgoing_index"):
            self._host[k] = None

    def _bake_pairs(self, vis, pairs, ff):
        """Per-pair device tables behind ``form_factors_tilde`` (RadiosityFast.py:403-433,
        :1234-1272) from the visibility matrix, the visible pairs and their form factors."""
        g = self._geom()
        dev = self._device
        with_brdf = self._brdf_incoming_directions is not None
        if with_brdf:
            vi, vo, brdf, bidx = self._brdf_tables()
            n_in, n_out = vi.shape[1], vo.shape[1]
            vi_d, vo_d = torch.from_numpy(vi).to(dev), torch.from_numpy(vo).to(dev)
        else:
            n_in = n_out = 1
            vi_d = vo_d = None
        dist, out_dir, in_dir = bake.pair_geometry(g["center"], g["wall_ids"], pairs,
                                                   vi_d, vo_d)
        n_bins = 1 if self._frequencies is None else self.n_bins
        # coef[c,d,b] = exp(-air[b] * 1.0) * brdf[c,d,b]: the reference evaluates
        # the air attenuation of the tilde at the norm of a *unit* vector
        # (RadiosityFast.py:1252-1262) and only when it is set at bake time
        air = (np.zeros(n_bins) if self._air_attenuation is None
               else np.real(self._air_attenuation).astype(float))
        if with_brdf:
            coef = np.exp(-air)[None, None, :] * brdf.reshape(-1, n_out, n_bins)
            sender_wall = g["wall_ids"][
                exchange.directed_pairs(pairs, ff, g["area"])[0]]
            cls_idx = torch.from_numpy(bidx).to(dev)[sender_wall] * n_in + in_dir.long()
        else:
            coef = np.exp(-air)[None, None, :] * np.ones((1, 1, n_bins))
            cls_idx = torch.zeros(2 * pairs.shape[0], dtype=torch.int64, device=dev)
        sender, receiver, ff_dir = exchange.directed_pairs(pairs, ff, g["area"])
        self._baked = dict(
            vis=vis, pairs=pairs, ff=ff, dist=dist, out_dir=out_dir, in_dir=in_dir,
            cls=cls_idx, coef=torch.from_numpy(np.ascontiguousarray(coef)).to(dev),
            sender=sender, receiver=receiver, ff_dir=ff_dir, with_brdf=with_brdf,
            n_out=n_out, n_bins=n_bins)

    # ------------------------------------------------------------------
    # source: RadiosityFast.py:436-522
    # ------------------------------------------------------------------
    def init_source_energy(self, source):
        """Initialize the source energy."""
        if isinstance(source, _COORD_TYPES):
            if source.cshape != (1, ):
                raise ValueError('just one source position is allowed.')
            source_position = np.asarray(source.cartesian, float).reshape(-1, 3)[0]
        elif isinstance(source, sound_object.SoundSource) or hasattr(source, "position"):
            source_position = np.asarray(source.position, float)
        else:
            raise ValueError(
                "source must be pf.Coordinates or sparrowpy SoundSource")
        self._source = source
        svis, d0, e0 = self._source_energy(source_position[None])
        self._source_vis_dev, self._d0_dev, self._e0_dev = svis[0], d0[0], e0[0]
        if getattr(source, "directivity", None) is not None:
            # one real factor per (patch, band), the same for every outgoing direction
            # (RadiosityFast.py:497-517); evaluated on the host, applied on the device
            fac = np.stack([np.real(source.get_directivity(self.patches_center, f))
                            for f in self._frequencies], axis=-1)
            self._e0_dev = self._e0_dev * torch.from_numpy(
                np.ascontiguousarray(fac, dtype=float)).to(self._device)[:, None, :]
        self._host["energy_init_source"] = None
        self._host["distance_patches_to_source"] = None

    def init_source_energy_batch(self, sources):
        """Extension: several source positions at once.

        ``sources``: Coordinates of cshape (S,).  The following
        ``calculate_energy_exchange`` propagates all S sources in the same kernel
        launches (a source is one more group of independent channels) and
        ``collect_energy_receiver_mono`` returns TimeData of cshape (S, R, n_bins).
        """
        if not isinstance(sources, _COORD_TYPES) or sources.cdim != 1:
            raise ValueError("sources must be pf.Coordinates of shape (n_sources, 3)")
        self._source = sources
        positions = np.asarray(sources.cartesian, float).reshape(-1, 3)
        self._source_vis_dev, self._d0_dev, self._e0_dev = self._source_energy(positions)
        self._host["energy_init_source"] = None
        self._host["distance_patches_to_source"] = None

    def _source_energy(self, positions):
        """Point visibility, source->patch energy and distances for S positions:
        (vis (S,N) bool, d0 (S,N), e0 (S,N,D,B)) on the device."""
        if self._brdf_incoming_directions is None:
            frequencies = np.array([0]) if self._frequencies is None else \
                self._frequencies
            self.set_wall_brdf(
                np.arange(self.n_walls),
                _FrequencyData(np.ones_like(frequencies, dtype=float), frequencies),
                _Coordinates(0, 0, 1, weights=1), _Coordinates(0, 0, 1, weights=1))
            self._frequencies = frequencies
        if self._air_attenuation is None:
            frequencies = np.array([0]) if self._frequencies is None else \
                self._frequencies
            self.set_air_attenuation(
                _FrequencyData(np.zeros_like(frequencies, dtype=float), frequencies))
            self._frequencies = frequencies

        g = self._geom()
        dev = self._device
        vi, vo, brdf, bidx = self._brdf_tables()
        src = torch.from_numpy(np.ascontiguousarray(positions, dtype=float)).to(dev)
        svis = bake.visibility_pt2p(src, g["center"], g["walls_normal"], g["walls_points"])
        air = torch.from_numpy(np.real(self._air_attenuation).astype(float)).to(dev)
        vi_d, brdf_d = torch.from_numpy(vi).to(dev), torch.from_numpy(brdf).to(dev)
        bidx_d = torch.from_numpy(bidx).to(dev)
        d0, e0 = bake.source_energy_batch(
            src, g["center"], g["points"], svis, air, g["wall_ids"], vi_d, brdf_d, bidx_d,
            vo.shape[1])
        return svis, d0, e0

    # ------------------------------------------------------------------
    # exchange: RadiosityFast.py:524-568
    # ------------------------------------------------------------------
    def _pair_tables(self, speed_of_sound, dt, n_samples, n_shards=1, shard=None):
        """Exchange tables for (c, dt, T); ``n_shards`` > 1 numbers the patches so that
        equal contiguous receiver shards are load-balanced (multi-GPU runs).  ``shard``
        (a rank index) keeps only the pairs received by that shard -- large scenes,
        where no rank can hold the tables of the whole scene."""
        key = (float(speed_of_sound), float(dt), int(n_samples), self._dtype, int(n_shards),
               shard)
        if self._tables is None or self._tables[0] != key:
            b = self._baked
            delay = bake.delay_bins(b["dist"], speed_of_sound, dt).long()
            delay = torch.stack([delay, delay], dim=1).reshape(-1)
            rank, n_internal = geometry.compact_patch_order(
                self._patches_points, self._patch_to_wall_ids, n_shards=n_shards)
            rng = None
            if shard is not None:
                from .distributed import shard_range
                lo, hi, _ = shard_range(n_internal, int(shard), int(n_shards))
                rng = (lo, hi)
            tables = exchange.build_pair_tables(
                b["sender"], b["receiver"], b["ff_dir"], delay, b["out_dir"], b["cls"],
                b["coef"], self.n_patches, n_samples, self._dtype,
                rank=torch.from_numpy(rank).to(self._device), n_internal=n_internal,
                receiver_range=rng)
            self._tables = (key, tables)
        return self._tables[1]

    def _resume(self):
        """Rebuild the device state that a restored checkpoint (``from_dict``) holds on the
        host only, so that the simulation continues at whatever stage it was saved
        (reference: tests/test_DirectionalRadiosityFast.py:72-125, ``cls(**input_dict)``
        RadiosityFast.py:882-886)."""
        h = self._host
        dev = self._device
        if (self._baked is None and h["visible_patches"] is not None
                and h["form_factors"] is not None):
            pairs = np.asarray(h["visible_patches"], dtype=np.int64).reshape(-1, 2)
            ffm = np.asarray(h["form_factors"], dtype=float)
            ff = np.ascontiguousarray(ffm[pairs[:, 0], pairs[:, 1]])
            vis = h["visibility_matrix"]
            if vis is None:
                vis = np.zeros((self.n_patches, self.n_patches), dtype=bool)
                vis[pairs[:, 0], pairs[:, 1]] = True
            keep = {k: h[k] for k in ("visibility_matrix", "visible_patches", "form_factors",
                                      "form_factors_tilde", "patch_2_brdf_outgoing_index")}
            self._bake_pairs(
                torch.from_numpy(np.ascontiguousarray(vis).astype(np.uint8)).to(dev),
                torch.from_numpy(pairs.astype(np.int32)).to(dev).contiguous(),
                torch.from_numpy(ff).to(dev))
            self._tables = None
            h.update(keep)                  # the restored arrays stay what was restored
        if (self._e0_dev is None and h["energy_init_source"] is not None
                and h["distance_patches_to_source"] is not None):
            self._e0_dev = torch.from_numpy(np.ascontiguousarray(
                h["energy_init_source"], dtype=float)).to(dev)
            self._d0_dev = torch.from_numpy(np.ascontiguousarray(
                h["distance_patches_to_source"], dtype=float)).to(dev)

    def calculate_energy_exchange(
            self, speed_of_sound, etc_time_resolution, etc_duration,
            max_reflection_order=-1, recalculate=False):
        """Calculate the energy exchange between patches."""
        n_samples = int(etc_duration / etc_time_resolution)
        have_result = self._hist is not None or \
            self._host["energy_exchange_etc"] is not None
        if not have_result or recalculate:
            self._resume()
            if self._e0_dev is None:
                raise _lib.SparrowB200Error(
                    "init_source_energy must be called before calculate_energy_exchange")
            delay0 = bake.delay_bins(self._d0_dev, speed_of_sound, etc_time_resolution)
            if max_reflection_order < 1:
                # initial energy only (RadiosityFast.py:550-555): no pair tables needed
                n, n_out, n_bins = self._e0_dev.shape[-3:]
                empty = torch.zeros(0, dtype=torch.int64, device=self._device)
                tables = exchange.build_pair_tables(
                    empty, empty, torch.zeros(0, dtype=torch.float64, device=self._device),
                    empty, empty, empty,
                    torch.ones((1, n_out, n_bins), dtype=torch.float64,
                               device=self._device), n, n_samples, self._dtype)
            else:
                if self._baked is None:
                    raise _lib.SparrowB200Error(
                        "bake_geometry must be called before calculate_energy_exchange")
                tables = self._pair_tables(speed_of_sound, etc_time_resolution, n_samples)
                if (tables.n_dirs, tables.n_bands) != tuple(self._e0_dev.shape[-2:]):
                    raise ValueError(
                        "BRDF / frequency layout changed after bake_geometry: "
                        f"baked (n_directions, n_bins)={(tables.n_dirs, tables.n_bands)}, "
                        f"source energy has {tuple(self._e0_dev.shape[-2:])}; "
                        "set the BRDF before bake_geometry")
            self._hist = exchange.energy_exchange(
                tables, self._e0_dev, delay0, n_samples, max_reflection_order)
            self._host["energy_exchange_etc"] = None
        self._etc_time_resolution = float(etc_time_resolution)
        self._speed_of_sound = float(speed_of_sound)
        self._etc_duration = float(etc_duration)

    # ------------------------------------------------------------------
    # receivers: RadiosityFast.py:570-752
    # ------------------------------------------------------------------
    def _receiver_tables(self, receiver_pos):
        g = self._geom()
        dev = self._device
        rcv = torch.from_numpy(np.atleast_2d(np.asarray(receiver_pos, float)).copy()).to(dev)
        hist = self._histogram()
        rvis = bake.visibility_pt2p(rcv, g["center"], g["walls_normal"], g["walls_points"])
        _, vo, _, _ = self._brdf_tables()
        air = torch.from_numpy(np.real(self._air_attenuation).astype(float)).to(dev)
        return hist, bake.receiver_factors(
            rcv, g["center"], g["points"], rvis, air, g["wall_ids"],
            torch.from_numpy(vo).to(dev), self.speed_of_sound, self._etc_time_resolution,
            hist.n_samples)

    def _histogram(self):
        if self._hist is None:
            etc = self._host["energy_exchange_etc"]
            if etc is None:
                raise _lib.SparrowB200Error(
                    "calculate_energy_exchange must be called first")
            # resume from a checkpointed ETC: upload into the padded layout
            n, d, b, t = etc.shape
            code = _lib.dtype_code(self._dtype)
            t_pad, pad = _lib.exchange_layout(t, 0, code)
            data = torch.zeros((b * n * d, t_pad + pad), dtype=_lib.torch_dtype(code),
                               device=self._device)
            band_major = np.ascontiguousarray(etc.transpose(2, 0, 1, 3))   # (B, N, D, T)
            data[:, pad:pad + t] = torch.from_numpy(band_major.reshape(b * n * d, t)).to(
                self._device)
            self._hist = exchange.EnergyHistogram(data, n, d, b, t, pad)
        return self._hist

    def collect_energy_receiver_mono(self, receivers, direct_sound=False):
        """Collect the energy at the receivers: TimeData of cshape (R, n_bins)."""
        if not isinstance(direct_sound, bool):
            raise ValueError("direct_sound must be of type boolean")
        self._check_receivers(receivers)
        hist, rt = self._receiver_tables(receivers.cartesian)
        mono = exchange.collect_mono(hist, rt["rdir"], rt["shift"], rt["scale"])
        etc_data = _np(mono.double())
        times = np.arange(etc_data.shape[-1]) * self._etc_time_resolution
        etc = _TimeData(etc_data, times)
        if direct_sound and hist.n_sources is not None:
            raise NotImplementedError("direct sound with a batch of sources")
        if direct_sound:
            direct, n_sample_delay = self.calculate_direct_sound(receivers)
            i_receivers = np.arange(len(n_sample_delay))
            etc.time[i_receivers, :, n_sample_delay] += direct
        return etc

    def collect_energy_receiver_patchwise(self, receivers):
        """Energy of every patch at the receivers: TimeData (R, n_patches, n_bins)."""
        self._check_receivers(receivers)
        hist, rt = self._receiver_tables(receivers.cartesian)
        out = exchange.collect_patchwise(hist, rt["rdir"], rt["shift"], rt["scale"])
        etc_data = _np(out.double())
        times = np.arange(etc_data.shape[-1]) * self._etc_time_resolution
        return _TimeData(etc_data, times)

    @staticmethod
    def _check_receivers(receivers):
        if not isinstance(receivers, _COORD_TYPES):
            raise ValueError("Receiver positions must be of type pf.Coordinates")
        if receivers.cdim != 1:
            raise ValueError("Receiver positions must be of shape (n_receivers, 3)")

    def calculate_direct_sound(self, receivers):
        """Direct sound (spreading loss, air attenuation) and its delay bins
        (RadiosityFast.py:605-657).  Host arithmetic: R x B values."""
        if not isinstance(receivers, _COORD_TYPES):
            raise ValueError("Receiver positions must be of type pf.Coordinates")
        if isinstance(self._source, _COORD_TYPES):
            source_position = np.asarray(self._source.cartesian, float).reshape(-1, 3)[0]
        else:
            source_position = np.asarray(self._source.position, float)
        diff = np.asarray(receivers.cartesian, float).reshape(-1, 3) - source_position
        r = np.sqrt(np.sum(diff ** 2, axis=-1))
        direct_sound = np.ones((r.shape[0], self.n_bins), dtype=float)
        direct_sound *= (1 / (4 * np.pi * r ** 2))[:, np.newaxis]
        if self._air_attenuation is not None:
            for i in range(self.n_bins):
                direct_sound[:, i] *= np.exp(-np.real(self._air_attenuation[i]) * r)
        if getattr(self._source, "directivity", None) is not None:
            rcv = np.asarray(receivers.cartesian, float).reshape(-1, 3)
            for i in range(self.n_bins):
                direct_sound[:, i] *= np.real(np.atleast_1d(self._source.get_directivity(
                    np.squeeze(rcv), self._frequencies[i])))
        n_sample_delay = np.array(
            r / self.speed_of_sound / self._etc_time_resolution, dtype=int)
        return direct_sound, n_sample_delay

    # ------------------------------------------------------------------
    # materials: RadiosityFast.py:754-826
    # ------------------------------------------------------------------
    def set_air_attenuation(self, air_attenuation):
        """Set air attenuation factor in Np/m (FrequencyData)."""
        self._check_set_frequency(air_attenuation.frequencies)
        self._air_attenuation = np.atleast_1d(np.asarray(air_attenuation.freq).squeeze())

    def set_wall_brdf(self, wall_indexes, brdf, incoming_directions,
                      outgoing_directions):
        """Set the wall BRDF representing scattering and absorption."""
        assert (np.asarray(incoming_directions.z) >= 0).all(), \
            "Sources must be in the positive half space"
        assert (np.asarray(outgoing_directions.z) >= 0).all(), \
            "Receivers must be in the positive half space"
        self._check_set_frequency(brdf.frequencies)
        if self._brdf_incoming_directions is None:
            self._brdf_incoming_directions = np.empty((self.n_walls), dtype=object)
            self._brdf_outgoing_directions = np.empty((self.n_walls), dtype=object)
            self._brdf_index = np.empty((self.n_walls), dtype=np.int64)
            self._brdf_index.fill(-1)
            self._brdf = []
        for i in np.atleast_1d(wall_indexes):
            incoming_rot, outgoing_rot = _rotate_coords_to_normal(
                self.walls_normal[i], self.walls_up_vector[i],
                incoming_directions, outgoing_directions)
            self._brdf_incoming_directions[i] = incoming_rot
            self._brdf_outgoing_directions[i] = outgoing_rot
        self._brdf.append(np.asarray(brdf.freq) * np.pi)
        self._brdf_index[np.atleast_1d(wall_indexes)] = len(self._brdf) - 1

    def _check_set_frequency(self, frequencies):
        frequencies = np.asarray(frequencies)
        if self._frequencies is None:
            self._frequencies = frequencies
        else:
            assert self._frequencies.size == frequencies.size, \
                "Number of frequency bins do not match"
            assert (self._frequencies == frequencies).all(), \
                "Frequencies do not match"

    # ------------------------------------------------------------------
    # numpy views of baked / computed state (materialised on access)
    # ------------------------------------------------------------------
    @property
    def _visibility_matrix(self):
        if self._host["visibility_matrix"] is None and self._baked is not None:
            self._host["visibility_matrix"] = _np(self._baked["vis"])
        return self._host["visibility_matrix"]

    @property
    def _visible_patches(self):
        if self._host["visible_patches"] is None and self._baked is not None:
            self._host["visible_patches"] = _np(self._baked["pairs"])
        return self._host["visible_patches"]

    @property
    def _form_factors(self):
        if self._host["form_factors"] is None and self._baked is not None:
            n = self.n_patches
            ffm = np.zeros((n, n))
            p = _np(self._baked["pairs"])
            ffm[p[:, 0], p[:, 1]] = _np(self._baked["ff"])
            self._host["form_factors"] = ffm
        return self._host["form_factors"]

    @property
    def _form_factors_tilde(self):
        """Dense (N, N, D, B) tensor -- built from the factored tables on demand."""
        if self._host["form_factors_tilde"] is None and self._baked is not None:
            b = self._baked
            n = self.n_patches
            coef = _np(b["coef"])
            tilde = np.zeros((n, n) + coef.shape[1:])
            s, r = _np(b["sender"]), _np(b["receiver"])
            tilde[s, r] = _np(b["ff_dir"])[:, None, None] * coef[_np(b["cls"])]
            self._host["form_factors_tilde"] = tilde
        return self._host["form_factors_tilde"]

    @property
    def _patch_2_brdf_outgoing_index(self):
        if self._host["patch_2_brdf_outgoing_index"] is None and self._baked is not None:
            b = self._baked
            n = self.n_patches
            if b["with_brdf"]:
                p2o = b["n_out"] * np.ones((n, n), dtype=np.int64)
                p2o[_np(b["sender"]), _np(b["receiver"])] = _np(b["out_dir"])
            else:
                p2o = np.zeros((n, n), dtype=np.int64)
            self._host["patch_2_brdf_outgoing_index"] = p2o
        return self._host["patch_2_brdf_outgoing_index"]

    @property
    def _energy_init_source(self):
        if self._host["energy_init_source"] is None and self._e0_dev is not None:
            self._host["energy_init_source"] = _np(self._e0_dev)
        return self._host["energy_init_source"]

    @property
    def _distance_patches_to_source(self):
        if self._host["distance_patches_to_source"] is None and self._d0_dev is not None:
            self._host["distance_patches_to_source"] = _np(self._d0_dev)
        return self._host["distance_patches_to_source"]

    @property
    def _source_visibility(self):
        return None if self._source_vis_dev is None else _np(self._source_vis_dev)

    @property
    def _energy_exchange_etc(self):
        if self._host["energy_exchange_etc"] is None and self._hist is not None:
            self._host["energy_exchange_etc"] = _np(self._hist.dense().double())
        return self._host["energy_exchange_etc"]

    @property
    def energy_exchange_etc_device(self):
        """(N, D, B, T) view of the ETC on the GPU (no host copy)."""
        return self._histogram().dense()

    # ------------------------------------------------------------------
    # checkpoint: RadiosityFast.py:828-886
    # ------------------------------------------------------------------
    def to_dict(self):
        """Convert the object to a dictionary (same keys as the reference)."""
        dict_out = {
            'walls_points': self._walls_points,
            'walls_normal': self._walls_normal,
            'walls_up_vector': self._walls_up_vector,
            'patches_points': self._patches_points,
            'n_patches': self._n_patches,
            'patch_to_wall_ids': self._patch_to_wall_ids,
            'visibility_matrix': self._visibility_matrix,
            'visible_patches': self._visible_patches,
            'form_factors': self._form_factors,
            'form_factors_tilde': self._form_factors_tilde,
            'frequencies': self._frequencies,
            'brdf': self._brdf,
            'brdf_index': self._brdf_index,
            'brdf_incoming_directions': self._brdf_incoming_directions,
            'brdf_outgoing_directions': self._brdf_outgoing_directions,
            'patch_2_brdf_outgoing_index': self._patch_2_brdf_outgoing_index,
            'air_attenuation': self._air_attenuation,
            'speed_of_sound': self._speed_of_sound,
            'etc_time_resolution': self._etc_time_resolution,
            'etc_duration': self._etc_duration,
            'distance_patches_to_source': self._distance_patches_to_source,
            'energy_init_source': self._energy_init_source,
            'energy_exchange_etc': self._energy_exchange_etc,
        }
        for key, value in dict_out.items():
            if value is None:
                dict_out[key] = 'None'
            elif isinstance(value, np.ndarray) and value.dtype != object:
                dict_out[key] = value.tolist()
        return dict_out

    def __eq__(self, other):
        """Equality of two objects = equality of their ``to_dict()`` (RadiosityFast.py:875-879;
        the reference uses deepdiff, here a numpy-aware recursive comparison)."""
        if not isinstance(other, DirectionalRadiosityFast):
            return False
        return _deep_equal(self.to_dict(), other.to_dict())

    __hash__ = None

    @classmethod
    def from_dict(cls, input_dict):
        """Create an object from a dictionary (resume from a checkpoint)."""
        data = {k: (None if isinstance(v, str) and v == 'None' else v)
                for k, v in input_dict.items()}
        return cls(**data)

    def write(self, filename, compress=True):
        """Write the object to a far file (needs pyfar)."""
        if _pf is None:
            raise ImportError("pyfar is required for .far files")
        _pf.io.write(filename, compress=compress, **self.to_dict())

    @classmethod
    def from_read(cls, filename):
        """Read the object from a far file (needs pyfar)."""
        if _pf is None:
            raise ImportError("pyfar is required for .far files")
        return cls.from_dict(_pf.io.read(filename))

    # ------------------------------------------------------------------
    # properties: RadiosityFast.py:888-968
    # ------------------------------------------------------------------
    @property
    def n_bins(self):
        return None if self._frequencies is None else self._frequencies.shape[0]

    @property
    def n_walls(self):
        return self._walls_points.shape[0]

    @property
    def n_patches(self):
        return self._n_patches

    @property
    def form_factors(self):
        return self._form_factors

    @property
    def visibility_matrix(self):
        return self._visibility_matrix

    @property
    def walls_area(self):
        return geometry.calculate_area(self._walls_points)

    @property
    def walls_points(self):
        return self._walls_points

    @property
    def walls_normal(self):
        return self._walls_normal

    @property
    def walls_center(self):
        return geometry.calculate_center(self._walls_points)

    @property
    def walls_up_vector(self):
        return self._walls_up_vector

    @property
    def patches_area(self):
        return geometry.calculate_area(self._patches_points)

    @property
    def patches_center(self):
        return geometry.calculate_center(self._patches_points)

    @property
    def patches_size(self):
        return geometry.calculate_size(self._patches_points)

    @property
    def patches_points(self):
        return self._patches_points

    @property
    def patches_normal(self):
        return self._walls_normal[self._patch_to_wall_ids]

    @property
    def speed_of_sound(self):
        return self._speed_of_sound


def _rotate_coords_to_normal(wall_normal, wall_up_vector, sources, receivers):
    """Rotate BRDF directions (frame: normal +z, up +x) into a wall's frame
    (RadiosityFast.py:971-986): the reference's pyfar calls restated on scipy's Rotation,
    see pyfar_shim.rotate_to_wall."""
    return (pyfar_shim.rotate_to_wall(sources, wall_normal, wall_up_vector),
            pyfar_shim.rotate_to_wall(receivers, wall_normal, wall_up_vector))


def _object_array(items):
    """BRDF direction lists as the reference keeps them: 1-D object arrays of Coordinates
    (RadiosityFast.py:789-792), whatever sequence a checkpoint handed in."""
    if items is None:
        return None
    out = np.empty(len(items), dtype=object)
    for i, item in enumerate(items):
        out[i] = item
    return out


def _deep_equal(a, b):
    """Recursive equality of checkpoint values: dicts, sequences, numpy arrays, Coordinates
    (compared by cartesian points and weights) and scalars."""
    if isinstance(a, dict) or isinstance(b, dict):
        return (isinstance(a, dict) and isinstance(b, dict) and a.keys() == b.keys()
                and all(_deep_equal(a[k], b[k]) for k in a))
    if isinstance(a, _COORD_TYPES) or isinstance(b, _COORD_TYPES):
        if not (isinstance(a, _COORD_TYPES) and isinstance(b, _COORD_TYPES)):
            return False
        wa, wb = getattr(a, "weights", None), getattr(b, "weights", None)
        return (_deep_equal(np.asarray(a.cartesian), np.asarray(b.cartesian))
                and _deep_equal(wa, wb))
    if isinstance(a, str) or isinstance(b, str):
        return isinstance(a, str) and isinstance(b, str) and a == b
    if a is None or b is None:
        return a is None and b is None
    seq = (list, tuple, np.ndarray)
    if isinstance(a, seq) or isinstance(b, seq):
        if not (isinstance(a, seq) and isinstance(b, seq)):
            return False
        aa = a if isinstance(a, np.ndarray) else None
        bb = b if isinstance(b, np.ndarray) else None
        if (aa is None or aa.dtype != object) and (bb is None or bb.dtype != object):
            try:
                aa, bb = np.asarray(a), np.asarray(b)
                if aa.dtype != object and bb.dtype != object:
                    return aa.shape == bb.shape and bool(np.array_equal(aa, bb))
            except ValueError:            # ragged: element by element below
                pass
        return len(a) == len(b) and all(_deep_equal(x, y) for x, y in zip(a, b))
    return bool(a == b)
