"""Procedural box terrains and env placement on them (BASELINE config 4, `Mjlab-Velocity-Rough-Unitree-G1`).

Restates, data-first (every sub-terrain is a function returning an array of axis-aligned boxes `[cx, cy, cz, hx, hy, hz]`
in tile coordinates plus a spawn origin), what the reference builds geom by geom through MuJoCo's spec API:

  terrains/terrain_generator.py:62-218   grid of tiles, random / curriculum difficulty, tile placement, outer border
  terrains/primitive_terrains.py:53-389  flat tile, pyramid stairs, inverted pyramid stairs
  terrains/utils.py:11-108               finite ground slab, four-box border
  terrains/terrain_importer.py:176-261   env origins: curriculum placement, level update, randomisation
  terrains/config.py:7-57                ROUGH_TERRAINS_CFG

Pinned by `tests/golden/terrain_rough.npz`, produced by running the reference's own generator against a recording stand-in
for `mujoco.MjSpec` (`tools/make_terrain_golden.py`).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _box(cx, cy, cz, sx, sy, sz):
  """Row for a box given its centre and FULL extents."""
  return [cx, cy, cz, sx / 2.0, sy / 2.0, sz / 2.0]


def ground_slab(size, height, thickness=1.0):
  """terrains/utils.py:11-35 (center_zero=False)."""
  return [_box(size[0] / 2.0, size[1] / 2.0, height - thickness / 2.0, size[0], size[1], thickness)]


def border_boxes(size, inner, height, center):
  """terrains/utils.py:38-108: top (+y), bottom (-y), left (-x), right (+x)."""
  tx, ty = (size[0] - inner[0]) / 2.0, (size[1] - inner[1]) / 2.0
  cx, cy, cz = center
  return [
    _box(cx, cy + inner[1] / 2.0 + ty / 2.0, cz, size[0], ty, height),
    _box(cx, cy - inner[1] / 2.0 - ty / 2.0, cz, size[0], ty, height),
    _box(cx - inner[0] / 2.0 - tx / 2.0, cy, cz, tx, inner[1], height),
    _box(cx + inner[0] / 2.0 + tx / 2.0, cy, cz, tx, inner[1], height),
  ]


@dataclass
class FlatTile:
  """BoxFlatTerrainCfg (primitive_terrains.py:53-64)."""

  proportion: float = 1.0

  def build(self, size, difficulty):
    return ground_slab(size, 0.0), np.array([size[0] / 2.0, size[1] / 2.0, 0.0])


@dataclass
class PyramidStairsTile:
  """BoxPyramidStairsTerrainCfg / BoxInvertedPyramidStairsTerrainCfg (primitive_terrains.py:67-389), `holes=False`."""

  proportion: float = 1.0
  step_height_range: tuple = (0.0, 0.1)
  step_width: float = 0.3
  platform_width: float = 1.0
  border_width: float = 0.0
  inverted: bool = False

  def build(self, size, difficulty):
    lo, hi = self.step_height_range
    sh = lo + difficulty * (hi - lo)
    sw, bw = self.step_width, self.border_width
    # floor division on floats, as the reference writes it (3.0 // 0.6 == 4.0)
    nx = (size[0] - 2 * bw - self.platform_width) // (2 * sw) + 1
    ny = (size[1] - 2 * bw - self.platform_width) // (2 * sw) + 1
    n = int(min(nx, ny))
    total = (n + 1) * sh
    cx, cy = 0.5 * size[0], 0.5 * size[1]
    tsx, tsy = size[0] - 2 * bw, size[1] - 2 * bw
    rows = []
    if bw > 0.0:
      rows += border_boxes(size, (tsx, tsy), sh, (cx, cy, -sh / 2))
    for k in range(n):
      bx, by = tsx - 2 * k * sw, tsy - 2 * k * sw
      off = (k + 0.5) * sw
      if self.inverted:
        z, h = -total / 2 - (k + 1) * sh / 2.0, total - (k + 1) * sh
      else:
        z, h = k * sh / 2.0, (k + 2) * sh
      rows += [
        _box(cx, cy + tsy / 2.0 - off, z, bx, sw, h),
        _box(cx, cy - tsy / 2.0 + off, z, bx, sw, h),
        _box(cx + tsx / 2.0 - off, cy, z, sw, by - 2 * sw, h),
        _box(cx - tsx / 2.0 + off, cy, z, sw, by - 2 * sw, h),
      ]
    px, py = tsx - 2 * n * sw, tsy - 2 * n * sw
    if self.inverted:
      rows.append(_box(cx, cy, -total - sh / 2, px, py, sh))
      origin = np.array([cx, cy, -(n + 1) * sh])
    else:
      rows.append(_box(cx, cy, n * sh / 2, px, py, (n + 2) * sh))
      origin = np.array([cx, cy, (n + 1) * sh])
    return rows, origin


@dataclass
class TerrainGeneratorCfg:
  """terrain_generator.py:48-59."""

  size: tuple = (8.0, 8.0)
  seed: int | None = 0
  curriculum: bool = False
  border_width: float = 0.0
  border_height: float = 1.0
  num_rows: int = 1
  num_cols: int = 1
  sub_terrains: dict = field(default_factory=dict)
  difficulty_range: tuple = (0.0, 1.0)


def rough_terrains_cfg(play: bool = False, seed: int | None = 0) -> TerrainGeneratorCfg:
  """ROUGH_TERRAINS_CFG (terrains/config.py:7-57) with the G1 rough task's overrides (config/g1/env_cfgs.py:54-55,143-148)."""
  stairs = dict(proportion=0.3, step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0, border_width=1.0)
  cfg = TerrainGeneratorCfg(size=(8.0, 8.0), seed=seed, border_width=20.0, num_rows=10, num_cols=20, curriculum=True,
                            sub_terrains={"flat": FlatTile(proportion=0.4), "pyramid_stairs": PyramidStairsTile(**stairs),
                                          "pyramid_stairs_inv": PyramidStairsTile(**stairs, inverted=True)})
  if play:
    cfg.curriculum, cfg.num_rows, cfg.num_cols, cfg.border_width = False, 5, 5, 10.0
  return cfg


@dataclass
class Terrain:
  """The generated grid: `boxes[n, 6]` (world centre, half extents; axis aligned) and `origins[rows, cols, 3]`."""

  boxes: np.ndarray
  origins: np.ndarray
  cfg: TerrainGeneratorCfg


def generate(cfg: TerrainGeneratorCfg) -> Terrain:
  """TerrainGenerator.__init__ + compile (terrain_generator.py:75-218); the RNG call sequence is the reference's."""
  if not cfg.sub_terrains:
    raise ValueError("At least one sub_terrain must be specified.")
  rng = np.random.default_rng(cfg.seed if cfg.seed is not None else np.random.randint(0, 10000))
  tiles = list(cfg.sub_terrains.values())
  prop = np.array([t.proportion for t in tiles], float)
  prop /= prop.sum()
  R, C = cfg.num_rows, cfg.num_cols
  origins = np.zeros((R, C, 3))
  boxes: list = []

  def place(r, c, tile, difficulty):
    corner = np.array([-R * cfg.size[0] * 0.5 + r * cfg.size[0], -C * cfg.size[1] * 0.5 + c * cfg.size[1], 0.0])
    rows, origin = tile.build(cfg.size, difficulty)
    rows = np.asarray(rows, float)
    rows[:, :3] += corner
    boxes.append(rows)
    origins[r, c] = origin + corner

  if cfg.curriculum:
    which = [int(np.min(np.where(c / C + 0.001 < np.cumsum(prop))[0])) for c in range(C)]
    lo, hi = cfg.difficulty_range
    for c in range(C):
      for r in range(R):
        place(r, c, tiles[which[c]], lo + (hi - lo) * ((r + rng.uniform()) / R))
  else:
    for i in range(R * C):
      r, c = divmod(i, C)
      k = rng.choice(len(prop), p=prop)
      place(r, c, tiles[k], rng.uniform(*cfg.difficulty_range))
  inner = (R * cfg.size[0], C * cfg.size[1])
  outer = (inner[0] + 2 * cfg.border_width, inner[1] + 2 * cfg.border_width)
  boxes.append(np.asarray(border_boxes(outer, inner, abs(cfg.border_height), (0, 0, -cfg.border_height / 2)), float))
  return Terrain(np.concatenate(boxes), origins, cfg)


# ---- env placement (TerrainImporter) ---------------------------------------------------------------------------------
def terrain_types(num_envs: int, num_cols: int) -> np.ndarray:
  """Column (terrain type) of each env: floor(env / (num_envs / num_cols)) (terrain_importer.py:236-240), evaluated the
  way `torch.div(..., rounding_mode="floor")` does in float32 (remainder first, then an exact quotient): with 64 envs and
  20 columns env 16 lands in column 4, not 5, because float32(3.2) > 3.2."""
  x, y = np.arange(num_envs, dtype=np.float32), np.float32(num_envs / num_cols)
  q = (x - np.fmod(x, y)) / y
  f = np.floor(q)
  return (f + (q - f > 0.5)).astype(np.int64)


def update_levels(levels, move_up, move_down, max_level, rand_level):
  """`update_env_origins` (terrain_importer.py:196-209): level += up - down; levels that reach `max_level` restart at
  `rand_level` (a uniform draw in [0, max_level)), negative ones clip to 0."""
  lv = levels + move_up.astype(levels.dtype) - move_down.astype(levels.dtype)
  return np.where(lv >= max_level, rand_level, np.clip(lv, 0, None))
