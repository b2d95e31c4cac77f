"""Rigid object descriptions (label, mesh, units).

Mirrors src/megapose/datasets/object_dataset.py:35-166 (`RigidObject`, `RigidObjectDataset`).
Symmetry handling (training losses / evaluation metrics) is out of scope of the inference hot path
and is carried only as plain attributes.  Extension: `mesh` may hold an in-memory
`megapose6d_b200.meshes.TriMesh` (procedural objects) instead of a file path.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Optional, Sequence, Set, Tuple


class RigidObject:
    def __init__(
        self,
        label: str,
        mesh_path: Optional[Path] = None,
        category: Optional[str] = None,
        mesh_diameter: Optional[float] = None,
        mesh_units: str = "m",
        symmetries_discrete: Sequence[Any] = (),
        symmetries_continuous: Sequence[Any] = (),
        ypr_offset_deg: Tuple[float, float, float] = (0.0, 0.0, 0.0),
        scaling_factor: float = 1.0,
        scaling_factor_mesh_units_to_meters: Optional[float] = None,
        mesh: Any = None,
    ):
        self.label = label
        self.category = category
        self.mesh_path = mesh_path
        self.mesh = mesh
        self.mesh_units = mesh_units
        if scaling_factor_mesh_units_to_meters is not None:
            self.scaling_factor_mesh_units_to_meters = scaling_factor_mesh_units_to_meters
        else:
            self.scaling_factor_mesh_units_to_meters = {"m": 1.0, "mm": 0.001}[mesh_units]
        self.scaling_factor = scaling_factor
        self.mesh_diameter = mesh_diameter
        self.diameter_meters = None
        self.symmetries_discrete = list(symmetries_discrete)
        self.symmetries_continuous = list(symmetries_continuous)
        self.ypr_offset_deg = ypr_offset_deg

    @property
    def is_symmetric(self) -> bool:
        return len(self.symmetries_discrete) > 0 or len(self.symmetries_continuous) > 0

    @property
    def scale(self) -> float:
        """Factor converting mesh coordinates to metres (object_dataset.py:120-123)."""
        return self.scaling_factor_mesh_units_to_meters * self.scaling_factor


class RigidObjectDataset:
    def __init__(self, objects: List[RigidObject]):
        self.list_objects = objects
        self.label_to_objects = {obj.label: obj for obj in objects}
        if len(self.list_objects) != len(self.label_to_objects):
            raise RuntimeError("There are objects with duplicate labels")

    def __getitem__(self, idx: int) -> RigidObject:
        return self.list_objects[idx]

    def __len__(self) -> int:
        return len(self.list_objects)

    def get_object_by_label(self, label: str) -> RigidObject:
        return self.label_to_objects[label]

    @property
    def objects(self) -> List[RigidObject]:
        return self.list_objects

    def filter_objects(self, keep_labels: Set[str]) -> "RigidObjectDataset":
        return RigidObjectDataset([o for o in self.list_objects if o.label in keep_labels])
