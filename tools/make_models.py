"""Compile the BASELINE robot scenes into committed model files.

Run in the build container (needs /root/reference for the robot MJCF, which is
input data, not code):

    python tools/make_models.py

Writes mjlab_b200/assets/{g1,go1}_{flat,rough,rough_play}.npz (rough: the ROUGH_TERRAINS_CFG box grid with the
fixed terrain seed 0 of SURVEY.md section 8(d), config 4).  Nothing at run time reads
/root/reference; the GPU box only sees the .npz files.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from mjlab_b200 import robots, terrains  # noqa: E402

REF = Path("/root/reference/src/mjlab/asset_zoo/robots")
SRC = {
  "g1": REF / "unitree_g1/xmls/g1.xml",
  "go1": REF / "unitree_go1/xmls/go1.xml",
}

if __name__ == "__main__":
  out = ROOT / "mjlab_b200" / "assets"
  out.mkdir(exist_ok=True)
  for which, xml in SRC.items():
    scene = robots.build_scene(xml.read_text(), which)
    m = scene.compile()
    m.save(out / f"{which}_flat.npz")
    print(which, dict(nq=m.nq, nv=m.nv, nu=m.nu, nbody=m.nbody, ngeom=m.ngeom, nsite=m.nsite,
                      nsensordata=m.nsensordata, npair=m.npair, meaninertia=round(m.stat_meaninertia, 5)))
  for name, play in (("g1_rough", False), ("g1_rough_play", True), ("go1_rough", False), ("go1_rough_play", True)):
    which = name.split("_")[0]
    ter = terrains.generate(terrains.rough_terrains_cfg(play=play, seed=0))
    m = robots.build_scene(SRC[which].read_text(), which, terrain=ter).compile()
    m.terrain_origins = np.asarray(ter.origins, float)
    m.terrain_size = np.asarray(ter.cfg.size, float)
    m.save(out / f"{name}.npz")
    print(name, dict(ngeom=m.ngeom, npair=m.npair, nfield=m.nfield, nfq=m.nfq, ndyn=m.ndyn, grid=(m.field_nx, m.field_ny),
                     items=len(m.field_cell_item)))
