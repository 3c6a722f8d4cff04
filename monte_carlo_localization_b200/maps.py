"""Occupancy-grid containers and map loading from ``maps/*.yaml`` (host-side helper).

The reference never decodes images itself: ``nav2_map_server`` (un-vendored, version
unpinned -- reference ``package.xml:34-35``, ``launch/mcl_launch.py:62-71``) serves an
``OccupancyGrid`` that ``get_omap`` consumes (``src/particle_filter.cpp:173-230``).  This
module restates nav2's published *trinary* conversion so the grids fed to the CUDA path and
to the CPU oracle are the same bytes:

    shade = mean(colour channels [+ alpha for trinary images with alpha]) / 255
    occ   = shade if negate else 1 - shade
    cell  = 100 if occ > occupied_thresh else 0 if occ < free_thresh else -1
    grid row 0 is the image's bottom row.

It is the Python twin of ``host/map_loader.cpp`` (the C++ host path); tests cross-check
the two.  ``resolution`` is kept as float32 because ``OccupancyGrid.info.resolution`` is
(``map_resolution_`` is that float widened to double, ``src/particle_filter.cpp:191``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class OccupancyGrid:
    """What ``get_omap`` reads from the map server: int8 row-major grid, row 0 = bottom."""

    data: np.ndarray          # int8 [height, width]; 0 free, 100 occupied, -1 unknown
    resolution: np.float32    # metres / cell (float32, as in the ROS message)
    origin: tuple             # (x, y, yaw) of cell (0, 0)'s lower-left corner
    name: str = ""

    @property
    def width(self) -> int:
        return int(self.data.shape[1])

    @property
    def height(self) -> int:
        return int(self.data.shape[0])

    @property
    def resolution_f64(self) -> float:
        return float(np.float64(np.float32(self.resolution)))

    def max_range_px(self, max_range: float = 12.0) -> int:
        # MAX_RANGE_PX = static_cast<int>(MAX_RANGE_METERS / map_resolution_)  (:195)
        return int(max_range / self.resolution_f64)

    def counts(self):
        d = self.data
        return int((d == 0).sum()), int((d == 100).sum()), int((d == -1).sum())


def _parse_map_yaml(path: str) -> dict:
    """Tiny YAML subset parser: scalars, flow lists ``[a, b, c]`` and block lists."""
    out: dict = {}
    key_for_list = None
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.split("#", 1)[0].rstrip()
            if not line.strip():
                continue
            s = line.strip()
            if s.startswith("- ") and key_for_list is not None:
                out[key_for_list].append(_scalar(s[2:].strip()))
                continue
            if ":" not in s:
                continue
            k, v = s.split(":", 1)
            k, v = k.strip(), v.strip()
            if v == "":
                out[k] = []
                key_for_list = k
                continue
            key_for_list = None
            if v.startswith("["):
                out[k] = [_scalar(t.strip()) for t in v.strip("[]").split(",") if t.strip()]
            else:
                out[k] = _scalar(v)
    return out


def _scalar(tok: str):
    tok = tok.strip().strip("'\"")
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def image_to_grid(img: np.ndarray, negate: bool, occupied_thresh: float, free_thresh: float,
                  has_alpha: bool = False) -> np.ndarray:
    """nav2 trinary conversion of an 8-bit image array [H, W] or [H, W, C] to int8 [H, W]."""
    a = np.asarray(img)
    if a.ndim == 2:
        chan_sum = a.astype(np.float64)
        nchan = 1
    else:
        c = a.shape[2]
        if c == 2:      # gray + alpha
            colour, alpha = a[..., :1], a[..., 1]
            has_alpha = True
        elif c == 4:
            colour, alpha = a[..., :3], a[..., 3]
            has_alpha = True
        else:
            colour, alpha = a[..., :3], None
            has_alpha = False
        # gray+alpha images: GraphicsMagick replicates gray into r,g,b
        if colour.shape[2] == 1:
            colour = np.repeat(colour, 3, axis=2)
        chan_sum = colour.astype(np.float64).sum(axis=2)
        nchan = 3
        if has_alpha and alpha is not None:
            chan_sum = chan_sum + alpha.astype(np.float64)
            nchan = 4
    if a.ndim == 2:
        # single-channel images read as r=g=b: mean of three equal quanta
        chan_sum = chan_sum * 3.0
        nchan = 3
    shade = (chan_sum / nchan) / 255.0
    occ = shade if negate else 1.0 - shade
    grid = np.full(occ.shape, -1, dtype=np.int8)
    grid[occ > occupied_thresh] = 100
    grid[(occ < free_thresh) & ~(occ > occupied_thresh)] = 0
    return np.ascontiguousarray(grid[::-1, :])  # row 0 = bottom of the image


def load_map_yaml(yaml_path: str) -> OccupancyGrid:
    """Load ``maps/<name>.yaml`` + its image the way ``nav2_map_server`` would serve it."""
    from PIL import Image

    meta = _parse_map_yaml(yaml_path)
    img_path = meta["image"]
    if not os.path.isabs(img_path):
        img_path = os.path.join(os.path.dirname(os.path.abspath(yaml_path)), img_path)
    im = Image.open(img_path)
    if im.mode in ("P", "1", "I;16", "I"):
        im = im.convert("L") if im.mode != "P" else im.convert("RGBA" if "transparency" in im.info else "RGB")
    arr = np.array(im)
    grid = image_to_grid(arr, bool(int(meta.get("negate", 0))), float(meta["occupied_thresh"]),
                         float(meta["free_thresh"]))
    org = meta["origin"]
    name = os.path.splitext(os.path.basename(yaml_path))[0]
    return OccupancyGrid(grid, np.float32(meta["resolution"]),
                         (float(org[0]), float(org[1]), float(org[2])), name)


def save_grid_npz(path: str, g: OccupancyGrid) -> None:
    np.savez_compressed(path, data=g.data, resolution=np.float32(g.resolution),
                        origin=np.asarray(g.origin, dtype=np.float64), name=np.asarray(g.name))


def load_grid_npz(path: str) -> OccupancyGrid:
    z = np.load(path)
    return OccupancyGrid(np.ascontiguousarray(z["data"].astype(np.int8)), np.float32(z["resolution"]),
                         tuple(float(v) for v in z["origin"]), str(z["name"]))


def fixture_path(name: str) -> str:
    """Decoded-grid fixtures committed under tests/golden/maps (made by tests/golden/make_fixtures.py)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(root, "tests", "golden", "maps", name + ".npz")


def load_named_map(name: str) -> OccupancyGrid:
    return load_grid_npz(fixture_path(name))


def synth_levine(width: int = 2048, height: int = 2048, seed: int = 7) -> OccupancyGrid:
    """Procedural stand-in for the missing ``levine.pgm`` honouring ``maps/levine.yaml``
    (res 0.05, origin -51.224998): a rectangular corridor loop with side rooms."""
    rng = np.random.default_rng(seed)
    g = np.full((height, width), -1, dtype=np.int8)
    cx0, cx1, cy0, cy1 = width // 2 - 420, width // 2 + 420, height // 2 - 300, height // 2 + 300
    half = 22  # 2.2 m wide corridor

    def carve(r0, r1, c0, c1):
        r0, r1, c0, c1 = max(r0, 1), min(r1, height - 1), max(c0, 1), min(c1, width - 1)
        g[r0 - 1:r1 + 1, c0 - 1:c1 + 1] = np.where(g[r0 - 1:r1 + 1, c0 - 1:c1 + 1] == 0, 0, 100)
        g[r0:r1, c0:c1] = 0

    carve(cy0 - half, cy0 + half, cx0 - half, cx1 + half)
    carve(cy1 - half, cy1 + half, cx0 - half, cx1 + half)
    carve(cy0 - half, cy1 + half, cx0 - half, cx0 + half)
    carve(cy0 - half, cy1 + half, cx1 - half, cx1 + half)
    for _ in range(24):  # side rooms off the corridors
        w, h = int(rng.integers(40, 90)), int(rng.integers(40, 90))
        if rng.random() < 0.5:
            c = int(rng.integers(cx0, cx1 - w))
            r = (cy0 - half - h + 2) if rng.random() < 0.5 else (cy1 + half - 2)
        else:
            r = int(rng.integers(cy0, cy1 - h))
            c = (cx0 - half - w + 2) if rng.random() < 0.5 else (cx1 + half - 2)
        carve(r, r + h, c, c + w)
    for _ in range(60):  # clutter
        r, c = int(rng.integers(cy0 - half, cy1 + half)), int(rng.integers(cx0 - half, cx1 + half))
        if g[r, c] == 0:
            g[r:r + 3, c:c + 3] = 100
    return OccupancyGrid(g, np.float32(0.05), (-51.224998, -51.224998, 0.0), "levine_synth")
