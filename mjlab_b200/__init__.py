"""mjlab_b200: B200-native world-batched physics step behind mjlab's `Simulation` API.

Only the hot path of SURVEY.md section 8 lives here: the model description
(`mjcf`, `robots`), the ctypes binding of the C ABI (`engine`), the
`Simulation` mirror (`sim`) and the CUDA sources (`csrc/`).
"""

from pathlib import Path

from .mjcf import ModelDesc, Spec  # noqa: F401

ASSETS = Path(__file__).resolve().parent / "assets"


def load_model(name: str) -> ModelDesc:
  """Load a committed compiled scene ('g1_flat', 'go1_flat')."""
  return ModelDesc.load(ASSETS / f"{name}.npz")
