"""Manager-term configuration objects of the reference that the B200 env layer consumes (src/mjlab/managers/).

`SceneEntityCfg` (managers/scene_entity_config.py:32-200): which joints / bodies / geoms / sites / actuators of a scene entity
a term acts on, by name (regex) or by local id; `resolve(scene)` makes names and ids consistent and collapses "everything, in
order" to `slice(None)`.
"""

from __future__ import annotations

from dataclasses import dataclass, field

_KINDS = (("joint", "find_joints", "num_joints"), ("body", "find_bodies", "num_bodies"), ("geom", "find_geoms", "num_geoms"),
          ("site", "find_sites", "num_sites"), ("actuator", "find_actuators", "num_actuators"))


def _all():
  return slice(None)


@dataclass
class SceneEntityCfg:
  name: str
  joint_names: str | tuple | list | None = None
  joint_ids: list | slice = field(default_factory=_all)
  body_names: str | tuple | list | None = None
  body_ids: list | slice = field(default_factory=_all)
  geom_names: str | tuple | list | None = None
  geom_ids: list | slice = field(default_factory=_all)
  site_names: str | tuple | list | None = None
  site_ids: list | slice = field(default_factory=_all)
  actuator_names: str | tuple | list | None = None
  actuator_ids: list | slice = field(default_factory=_all)
  preserve_order: bool = False

  def resolve(self, scene) -> None:
    entity = scene[self.name]
    for kind, find, num in _KINDS:
      names, ids = getattr(self, f"{kind}_names"), getattr(self, f"{kind}_ids")
      if names is None and not isinstance(ids, (int, list)):
        continue
      every = list(getattr(entity, f"{kind}_names"))
      if isinstance(names, str):
        names = [names]
      elif names is not None:
        names = list(names)
      if isinstance(ids, int):
        ids = [ids]
      if names is not None and isinstance(ids, list):
        found, _ = getattr(entity, find)(names, preserve_order=self.preserve_order)
        have = [every[i] for i in ids]
        if found != ids or have != names:
          raise ValueError(f"Inconsistent {kind} names and indices. Names {names} resolved to indices {found}, "
                           f"but indices {ids} (mapping to names {have}) were provided.")
      elif names is not None:
        found, _ = getattr(entity, find)(names, preserve_order=self.preserve_order)
        ids = slice(None) if (len(found) == getattr(entity, num) and names == every) else found
      else:
        names = [every[i] for i in ids]
      setattr(self, f"{kind}_names", names)
      setattr(self, f"{kind}_ids", ids)
