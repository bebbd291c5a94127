"""`Pointclouds`: a ragged batch of 3-D point clouds with named per-point features.

API-compatible with the reference container (structures/point_structure.py:40-1420): the same
constructor (`points` = list of (P_n,3) tensors or padded (N,P,3) tensor; `features` = dict
name -> list | padded tensor), the same list / padded / packed accessors, auxiliary index
tensors, batch operations (`__getitem__`, `clone`, `detach`, `to`, `extend`, `split`,
`offset_`, `scale_`, `update_padded`, `inside_box`) and module-level helpers
(`join_pointclouds_as_batch`, `join_pointclouds_as_scene`, `get_bounding_boxes`, `offset`,
`scale`, `subsample`, `all_close`).

What differs is the plumbing on the hot path: list -> padded runs as one
ragged-copy CUDA kernel per tensor (see structures/utils.py) rather than N slice assignments,
and `padded_to_packed_idx` is vectorised instead of a Python loop of per-cloud `arange`s.
"""
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import utils as struct_utils

Device = Union[str, torch.device]


def make_device(device: Device) -> torch.device:
    """str | torch.device -> torch.device; a bare "cuda" is pinned to the current device."""
    device = torch.device(device) if isinstance(device, str) else device
    if device.type == "cuda" and device.index is None:
        device = torch.device(f"cuda:{torch.cuda.current_device()}")
    return device


class Pointclouds:
    """Batch of N point clouds in three interchangeable layouts:

    * list   -- N tensors (P_n, 3) (+ per feature N tensors (P_n, C));
    * padded -- (N, max P_n, 3), zero padded;
    * packed -- (sum P_n, 3), with `packed_to_cloud_idx`, `cloud_to_packed_first_idx`,
      `num_points_per_cloud` and `padded_to_packed_idx` to move between them.

    Layouts are computed lazily from whichever one was supplied and cached.
    """

    _INTERNAL_TENSORS = [
        "_points_packed",
        "_points_padded",
        "_features_packed",
        "_features_padded",
        "_packed_to_cloud_idx",
        "_cloud_to_packed_first_idx",
        "_num_points_per_cloud",
        "_padded_to_packed_idx",
        "valid",
        "equisized",
    ]

    # ------------------------------------------------------------------ construction
    def __init__(self, points, features=None) -> None:
        self.device = torch.device("cpu")
        self.equisized = False
        self.valid = None
        self._N = 0
        self._P = 0
        self._C: Dict[str, int] = {}
        self._points_list = None
        self._features_list: Dict[str, List[torch.Tensor]] = {}
        self._num_points_per_cloud = None
        self._points_packed = None
        self._features_packed: Dict[str, torch.Tensor] = {}
        self._packed_to_cloud_idx = None
        self._cloud_to_packed_first_idx = None
        self._points_padded = None
        self._features_padded: Dict[str, torch.Tensor] = {}
        self._padded_to_packed_idx = None

        if isinstance(points, list):
            self._init_from_list(points)
        elif torch.is_tensor(points):
            if points.dim() != 3 or points.shape[2] != 3:
                raise ValueError("Points tensor has incorrect dimensions.")
            self._points_padded = points
            self._N, self._P = points.shape[0], points.shape[1]
            self.device = points.device
            self.valid = torch.ones((self._N,), dtype=torch.bool, device=self.device)
            self._num_points_per_cloud = torch.tensor([self._P] * self._N, device=self.device)
            self.equisized = True
        else:
            raise ValueError(
                "Points must be either a list or a tensor with \
                    shape (batch_size, P, 3) where P is the maximum number of \
                    points in a cloud."
            )

        if features is not None:
            if not isinstance(features, dict):
                raise ValueError("Features must be a dictionary with feature names as keys")
            for name, data in features.items():
                as_list, as_padded, channels = self._parse_auxiliary_input(data)
                if as_list is not None:
                    self._features_list[name] = as_list
                elif as_padded is not None:
                    self._features_padded[name] = as_padded
                else:
                    raise ValueError(
                        "Features must be either a list or a padded tensor with \
                            shape (batch_size, P, C) where P is the maximum number of \
                            points in a cloud and C is the number of channels."
                    )
                self._C[name] = channels if channels is not None else 0

    def _init_from_list(self, points: list) -> None:
        self._points_list = points
        self._N = len(points)
        self.valid = torch.zeros((self._N,), dtype=torch.bool, device=self.device)
        if self._N == 0:
            self._num_points_per_cloud = torch.tensor([], dtype=torch.int64)
            return
        self.device = points[0].device
        for p in points:
            if len(p) > 0 and (p.dim() != 2 or p.shape[1] != 3):
                raise ValueError("Clouds in list must be of shape Px3 or empty")
            if p.device != self.device:
                raise ValueError("All points must be on the same device")
        sizes = [len(p) for p in points]  # host ints: no device sync for max / unique
        self._num_points_per_cloud = torch.tensor(sizes, device=self.device)
        self._P = max(sizes)
        self.valid = torch.tensor([s > 0 for s in sizes], dtype=torch.bool, device=self.device)
        self.equisized = len(set(sizes)) == 1

    def _parse_auxiliary_input(
        self, aux_input
    ) -> Tuple[Optional[List[torch.Tensor]], Optional[torch.Tensor], Optional[int]]:
        """features value -> (list, padded, C); exactly one of list / padded is not None."""
        if aux_input is None or self._N == 0:
            return None, None, None
        if isinstance(aux_input, list):
            return self._parse_auxiliary_input_list(aux_input)
        if torch.is_tensor(aux_input):
            if aux_input.dim() != 3:
                raise ValueError("Auxiliary input tensor has incorrect dimensions.")
            if self._N != aux_input.shape[0]:
                raise ValueError("Points and inputs must be the same length.")
            if self._P != aux_input.shape[1]:
                raise ValueError(
                    "Inputs tensor must have the right maximum \
                    number of points in each cloud."
                )
            if aux_input.device != self.device:
                raise ValueError("All auxiliary inputs must be on the same device as the points.")
            return None, aux_input, aux_input.shape[2]
        raise ValueError(
            "Auxiliary input must be either a list or a tensor with \
                    shape (batch_size, P, C) where P is the maximum number of \
                    points in a cloud."
        )

    def _parse_auxiliary_input_list(
        self, aux_input: list
    ) -> Tuple[Optional[List[torch.Tensor]], None, Optional[int]]:
        """List form of a feature: validate per cloud, replace malformed empties by (0, C)."""
        if len(aux_input) != self._N:
            raise ValueError("Points and auxiliary input must be the same length.")
        sizes = self._num_points_per_cloud.tolist()
        channels = None
        usable = []
        for p, d in zip(sizes, aux_input):
            ok = p > 0 or (d is not None and d.ndim == 2)
            usable.append(ok)
            if not ok:
                continue
            if p != d.shape[0]:
                raise ValueError("A cloud has mismatched numbers of points and inputs")
            if d.dim() != 2:
                raise ValueError("A cloud auxiliary input must be of shape PxC or empty")
            if channels is None:
                channels = d.shape[1]
            elif channels != d.shape[1]:
                raise ValueError("The clouds must have the same number of channels")
            if d.device != self.device:
                raise ValueError("All auxiliary inputs must be on the same device as the points.")
        if channels is None:
            return None, None, None
        if all(usable):
            return aux_input, None, channels
        empty = torch.zeros((0, channels), device=self.device)
        return [d if ok else empty for ok, d in zip(usable, aux_input)], None, channels

    # ------------------------------------------------------------------ list accessors
    def points_list(self) -> List[torch.Tensor]:
        """List of (P_n, 3) tensors (views into the padded tensor when built from one)."""
        if self._points_list is None:
            assert self._points_padded is not None, "points_padded is required to compute points_list."
            sizes = self.num_points_per_cloud().tolist()
            self._points_list = [self._points_padded[i, :s] for i, s in enumerate(sizes)]
        return self._points_list

    def get_features_list(self, feature_name: str) -> Optional[List[torch.Tensor]]:
        """List of (P_n, C) tensors of one feature, or None if absent."""
        if feature_name not in self._features_list:
            if feature_name not in self._features_padded:
                return None
            self._features_list[feature_name] = struct_utils.padded_to_list(
                self._features_padded[feature_name], self.num_points_per_cloud().tolist()
            )
        return self._features_list[feature_name]

    def features_list(self) -> Dict[str, List[torch.Tensor]]:
        """name -> list of (P_n, C) tensors, for every feature."""
        out = {}
        for name in set(self._features_list) | set(self._features_padded):
            as_list = self.get_features_list(name)
            if as_list is not None:
                out[name] = as_list
        return out

    # ------------------------------------------------------------------ packed accessors
    def _compute_packed(self, refresh: bool = False):
        """Build the packed tensors and their index helpers from the list layout."""
        have = (self._points_packed, self._packed_to_cloud_idx, self._cloud_to_packed_first_idx)
        if not refresh and all(v is not None for v in have):
            return
        points_list = self.points_list()
        features = self.features_list()
        if self.isempty():
            self._points_packed = torch.zeros((0, 3), dtype=torch.float32, device=self.device)
            self._packed_to_cloud_idx = torch.zeros((0,), dtype=torch.int64, device=self.device)
            self._cloud_to_packed_first_idx = torch.zeros((0,), dtype=torch.int64, device=self.device)
            self._features_packed = {}
            return
        packed, counts, first, to_cloud = struct_utils.list_to_packed(points_list)
        if not torch.allclose(self._num_points_per_cloud, counts):
            raise ValueError("Inconsistent list to packed conversion")
        self._points_packed = packed
        self._cloud_to_packed_first_idx = first
        self._packed_to_cloud_idx = to_cloud
        self._features_packed = {
            name: torch.cat(as_list, dim=0) for name, as_list in features.items() if as_list is not None
        }

    def points_packed(self) -> torch.Tensor:
        """(sum P_n, 3)."""
        self._compute_packed()
        return self._points_packed

    def get_features_packed(self, feature_name: str) -> Optional[torch.Tensor]:
        """(sum P_n, C) of one feature, or None if absent."""
        self._compute_packed()
        return self._features_packed.get(feature_name)

    def features_packed(self) -> Dict[str, torch.Tensor]:
        """name -> (sum P_n, C)."""
        self._compute_packed()
        return self._features_packed

    # ------------------------------------------------------------------ padded accessors
    def _compute_padded(self, refresh: bool = False):
        """Build the padded tensors from the list layout (one ragged-copy kernel per tensor on
        CUDA, see structures/utils.list_to_padded)."""
        if not refresh and self._points_padded is not None:
            return
        self._features_padded = {}
        if self.isempty():
            self._points_padded = torch.zeros((self._N, 0, 3), device=self.device)
            return
        self._points_padded = struct_utils.list_to_padded(
            self.points_list(), (self._P, 3), pad_value=0.0, equisized=self.equisized
        )
        for name, as_list in self.features_list().items():
            if as_list is None or len(as_list) == 0:
                continue
            channels = as_list[0].shape[1] if as_list[0].dim() > 1 else 1
            self._features_padded[name] = struct_utils.list_to_padded(
                as_list, (self._P, channels), pad_value=0.0, equisized=self.equisized
            )

    def points_padded(self) -> torch.Tensor:
        """(N, max P_n, 3), zero padded."""
        self._compute_padded()
        return self._points_padded

    def get_features_padded(self, feature_name: str) -> Optional[torch.Tensor]:
        """(N, max P_n, C) of one feature, or None if absent."""
        self._compute_padded()
        return self._features_padded.get(feature_name)

    def features_padded(self) -> Dict[str, torch.Tensor]:
        """name -> (N, max P_n, C)."""
        self._compute_padded()
        return self._features_padded

    # ------------------------------------------------------------------ index helpers
    def num_points_per_cloud(self) -> torch.Tensor:
        """(N,) number of points of each cloud."""
        return self._num_points_per_cloud

    def packed_to_cloud_idx(self):
        """(sum P_n,) cloud index of every packed point."""
        self._compute_packed()
        return self._packed_to_cloud_idx

    def cloud_to_packed_first_idx(self):
        """(N,) index of each cloud's first point in the packed layout."""
        self._compute_packed()
        return self._cloud_to_packed_first_idx

    def padded_to_packed_idx(self):
        """(sum P_n,) indices such that points_padded().reshape(-1, 3)[idx] == points_packed()."""
        if self._padded_to_packed_idx is None:
            if self._N == 0:
                self._padded_to_packed_idx = []
            else:
                self._padded_to_packed_idx = struct_utils.padded_to_packed_index(
                    self.num_points_per_cloud().to(torch.int64), self._P
                )
        return self._padded_to_packed_idx

    # ------------------------------------------------------------------ batch operations
    def __len__(self) -> int:
        return self._N

    def __getitem__(
        self, index: Union[int, List[int], slice, torch.BoolTensor, torch.LongTensor]
    ) -> "Pointclouds":
        """Sub-batch (tensors are shared, not cloned).  int, slice, list of ints, bool / long tensor."""
        if isinstance(index, int):
            pick = lambda seq: [seq[index]]  # noqa: E731
        elif isinstance(index, slice):
            pick = lambda seq: seq[index]  # noqa: E731
        elif isinstance(index, list):
            pick = lambda seq: [seq[i] for i in index]  # noqa: E731
        elif isinstance(index, torch.Tensor):
            if index.dim() != 1 or index.dtype.is_floating_point:
                raise IndexError(index)
            if index.dtype == torch.bool:
                index = index.nonzero()
                index = index.squeeze(1) if index.numel() > 0 else index
                index = index.tolist()
            pick = lambda seq: [seq[i] for i in index]  # noqa: E731
        else:
            raise IndexError(index)
        features = {name: pick(as_list) for name, as_list in self.features_list().items()}
        return self.__class__(points=pick(self.points_list()), features=features if features else None)

    def isempty(self) -> bool:
        """True when there is no cloud, or every cloud has zero points."""
        return self._N == 0 or self.valid.eq(False).all()

    def _rebuild(self, fn):
        """New Pointclouds whose source tensors and cached internals are `fn(tensor)`."""
        points, features = None, None
        if self._points_list is not None:
            points = [fn(v) for v in self.points_list()]
            as_lists = self.features_list()
            if as_lists:
                features = {name: [fn(f) for f in lst] for name, lst in as_lists.items()}
        elif self._points_padded is not None:
            points = fn(self.points_padded())
            as_padded = self.features_padded()
            if as_padded:
                features = {name: fn(t) for name, t in as_padded.items()}
        other = self.__class__(points=points, features=features)
        for k in self._INTERNAL_TENSORS:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(other, k, fn(v))
            elif isinstance(v, dict):
                setattr(other, k, {key: fn(val) if torch.is_tensor(val) else val for key, val in v.items()})
        return other

    def clone(self):
        """Deep copy (every tensor cloned)."""
        return self._rebuild(lambda t: t.clone())

    def detach(self):
        """Copy sharing storage with every tensor detached from autograd."""
        return self._rebuild(lambda t: t.detach())

    def to(self, device: Device, copy: bool = False):
        """torch.Tensor.to semantics: self when already there and copy is False."""
        device_ = make_device(device)
        if not copy and self.device == device_:
            return self
        other = self.clone()
        if self.device == device_:
            return other
        other.device = device_
        if other._N > 0:
            other._points_list = [v.to(device_) for v in other.points_list()]
            for name, as_list in other.features_list().items():
                other._features_list[name] = [f.to(device_) for f in as_list]
        for k in self._INTERNAL_TENSORS:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(other, k, v.to(device_))
            elif isinstance(v, dict):
                setattr(other, k, {key: val.to(device_) if torch.is_tensor(val) else val for key, val in v.items()})
        return other

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self.to("cuda")

    def extend(self, N: int):
        """Batch with every cloud repeated N times (cloned)."""
        if not isinstance(N, int):
            raise ValueError("N must be an integer.")
        if N <= 0:
            raise ValueError("N must be > 0.")
        points = [p.clone() for p in self.points_list() for _ in range(N)]
        features = {
            name: [f.clone() for f in as_list for _ in range(N)]
            for name, as_list in self.features_list().items()
        }
        return self.__class__(points=points, features=features)

    def split(self, split_sizes: list):
        """List of sub-batches of the given sizes (like torch.split)."""
        if not all(isinstance(x, int) for x in split_sizes):
            raise ValueError("Value of split_sizes must be a list of integers.")
        out, start = [], 0
        for size in split_sizes:
            out.append(self[start : start + size])
            start += size
        return out

    def get_cloud(self, index: int):
        """(points (P,3), {name: (P,C)}) of one cloud."""
        if not isinstance(index, int):
            raise ValueError("Cloud index must be an integer.")
        if index < 0 or index > self._N:
            raise ValueError(
                "Cloud index must be in the range [0, N) where \
            N is the number of clouds in the batch."
            )
        features = {
            name: as_list[index] for name, as_list in self.features_list().items() if as_list is not None
        }
        return self.points_list()[index], features

    # ------------------------------------------------------------------ in-place geometry
    def _refresh_padded_points(self, new_points_list) -> None:
        if self._points_padded is not None:
            for i, pts in enumerate(new_points_list):
                if len(pts) > 0:
                    self._points_padded[i, : pts.shape[0], :] = pts

    def offset_(self, offsets_packed):
        """Add offsets ((3,) or (sum P_n, 3)) to every point, in place.  Returns self."""
        packed = self.points_packed()
        if offsets_packed.shape == (3,):
            offsets_packed = offsets_packed.expand_as(packed)
        if offsets_packed.shape != packed.shape:
            raise ValueError("Offsets must have dimension (all_p, 3).")
        self._points_packed = packed + offsets_packed
        self._points_list = list(self._points_packed.split(self.num_points_per_cloud().tolist(), 0))
        self._refresh_padded_points(self._points_list)
        return self

    def scale_(self, scale):
        """Multiply coordinates by a scalar or a per-cloud (N,) tensor, in place.  Returns self."""
        if not torch.is_tensor(scale):
            scale = torch.full((len(self),), scale, device=self.device)
        self._points_list = [scale[i] * pts for i, pts in enumerate(self.points_list())]
        if self._points_packed is not None:
            self._points_packed = torch.cat(self._points_list, dim=0)
        self._refresh_padded_points(self._points_list)
        return self

    def update_padded(self, new_points_padded, new_features_padded=None):
        """New Pointclouds with replaced padded points (and optionally features), sharing the
        index helpers; features are kept when none are given."""

        def check(x, size):
            if x.shape[0] != size[0]:
                raise ValueError("new values must have the same batch dimension.")
            if x.shape[1] != size[1]:
                raise ValueError("new values must have the same number of points.")
            if size[2] is not None and x.shape[2] != size[2]:
                raise ValueError("new values must have the same number of channels.")

        check(new_points_padded, [self._N, self._P, 3])
        if new_features_padded is not None:
            if not isinstance(new_features_padded, dict):
                raise ValueError("new_features_padded must be a dictionary")
            for name, t in new_features_padded.items():
                check(t, [self._N, self._P, self._C[name]])

        new = self.__class__(points=new_points_padded, features=new_features_padded)
        new.equisized = self.equisized
        if new_features_padded is None:
            new._features_list = self._features_list
            new._features_padded = self._features_padded
            new._features_packed = self._features_packed
        for k in ("_packed_to_cloud_idx", "_cloud_to_packed_first_idx", "_num_points_per_cloud",
                  "_padded_to_packed_idx", "valid"):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(new, k, v)
        new._points_padded = new_points_padded
        assert new._points_list is None
        assert new._points_packed is None
        if new_features_padded is not None:
            new._features_padded = new_features_padded
            new._features_list = {}
            new._features_packed = {}
        return new

    def inside_box(self, box):
        """Bool (sum P_n,) mask of packed points inside box ((2,3) or (N,2,3): [min; max])."""
        if box.dim() > 3 or box.dim() < 2:
            raise ValueError("Input box must be of shape (2, 3) or (N, 2, 3).")
        if box.dim() == 3 and box.shape[0] != 1 and box.shape[0] != self._N:
            raise ValueError("Input box dimension is incompatible with pointcloud size.")
        if box.dim() == 2:
            box = box[None]
        if (box[..., 0, :] > box[..., 1, :]).any():
            raise ValueError("Input box is invalid: min values larger than max values.")
        packed = self.points_packed()
        if box.shape[0] == 1:
            box = box.expand(packed.shape[0], 2, 3)
        elif box.shape[0] == self._N:
            box = box[self.packed_to_cloud_idx()]
        inside = (packed >= box[:, 0]) * (packed <= box[:, 1])
        return inside.all(dim=-1)


# ---------------------------------------------------------------------- module-level helpers
def join_pointclouds_as_batch(pointclouds: Sequence[Pointclouds]) -> Pointclouds:
    """Concatenate several Pointclouds into one batch; a feature survives only if every input
    carries it (with equal channel counts)."""
    if isinstance(pointclouds, Pointclouds) or not isinstance(pointclouds, Sequence):
        raise ValueError("Wrong first argument to join_points_as_batch.")
    device = pointclouds[0].device
    if not all(p.device == device for p in pointclouds):
        raise ValueError("Pointclouds must all be on the same device")
    per_cloud = [p.points_list() for p in pointclouds]
    if None in per_cloud:
        raise ValueError("Pointclouds cannot have their points set to None!")
    points = [p for lst in per_cloud for p in lst]
    feature_dicts = [p.features_list() for p in pointclouds]
    names = set().union(*[d.keys() for d in feature_dicts]) if feature_dicts else set()
    combined = {}
    for name in names:
        if not all(name in d and d[name] is not None for d in feature_dicts):
            continue
        merged = [f for d in feature_dicts for f in d[name]]
        if len(merged) > 0 and any(f.shape[1] != merged[0].shape[1] for f in merged[1:]):
            raise ValueError(
                f"Pointclouds must have the same number of channels for feature '{name}'"
            )
        combined[name] = merged
    return Pointclouds(points=points, features=combined if combined else None)


def join_pointclouds_as_scene(pointclouds: Union[Pointclouds, List[Pointclouds]]) -> Pointclouds:
    """Merge a batch (or a list of batches) into a single cloud."""
    if isinstance(pointclouds, list):
        pointclouds = join_pointclouds_as_batch(pointclouds)
    if len(pointclouds) == 1:
        return pointclouds
    features = {name: t[None] for name, t in pointclouds.features_packed().items()}
    return Pointclouds(points=pointclouds.points_packed()[None], features=features if features else None)


def get_bounding_boxes(pointcloud: "Pointclouds") -> torch.Tensor:
    """(N, 3, 2): per cloud, per axis, [min, max]."""
    mins = torch.stack([p.min(dim=0)[0] for p in pointcloud.points_list()], dim=0)
    maxs = torch.stack([p.max(dim=0)[0] for p in pointcloud.points_list()], dim=0)
    return torch.stack([mins, maxs], dim=2)


def offset(pointcloud: "Pointclouds", offsets_packed: torch.Tensor) -> "Pointclouds":
    """Out-of-place `offset_`."""
    return pointcloud.clone().offset_(offsets_packed)


def scale(pointcloud: "Pointclouds", scale: Union[float, torch.Tensor]) -> "Pointclouds":
    """Out-of-place `scale_`."""
    return pointcloud.clone().scale_(scale)


def subsample(pointclouds: Pointclouds, max_points: Union[int, Sequence[int]]) -> "Pointclouds":
    """Randomly keep at most max_points points of each cloud (features follow); returns the
    input unchanged when nothing exceeds the limit."""
    if isinstance(max_points, int):
        max_points = [max_points] * len(pointclouds)
    elif len(max_points) != len(pointclouds):
        raise ValueError("wrong number of max_points supplied")
    sizes = [int(s) for s in pointclouds.num_points_per_cloud()]
    limits = [int(m) for m in max_points]
    if all(s <= m for s, m in zip(sizes, limits)):
        return pointclouds
    features_in = pointclouds.features_list()
    points_out = []
    features_out = {name: [] for name in features_in}
    for i, (limit, size, pts) in enumerate(zip(limits, sizes, pointclouds.points_list())):
        keep = None
        if size > limit:
            keep = torch.tensor(np.random.choice(size, limit, replace=False), device=pts.device,
                                dtype=torch.int64)
            pts = pts[keep]
        for name, as_list in features_in.items():
            if as_list is None or i >= len(as_list):
                features_out[name].append(None)
            else:
                features_out[name].append(as_list[i] if keep is None else as_list[i][keep])
        points_out.append(pts)
    features_out = {n: lst for n, lst in features_out.items() if any(f is not None for f in lst)}
    return Pointclouds(points=points_out, features=features_out if features_out else None)


def all_close(pcd1: Pointclouds, pcd2: Pointclouds, rtol=1e-05, atol=1e-08, verbose=False) -> bool:
    """True when packed points and every packed feature agree within tolerance."""
    if pcd1.device != pcd2.device:
        raise ValueError("Pointclouds must be on the same device.")
    points_ok = torch.allclose(pcd1.points_packed(), pcd2.points_packed(), rtol, atol)
    if verbose:
        print("Points all close:", points_ok)
    keys1, keys2 = pcd1.features_packed().keys(), pcd2.features_packed().keys()
    if set(keys1) != set(keys2):
        if verbose:  # the key views themselves, as the reference prints them (:1399-1406)
            print("Features keys mismatch:", "Keys in pcd1:", keys1, "Keys in pcd2:", keys2)
        return False
    feats_ok = {
        name: torch.allclose(pcd1.get_features_packed(name), pcd2.get_features_packed(name), rtol, atol)
        for name in keys1
    }
    if verbose:
        print("Features all close:", feats_ok)
    return points_ok and all(feats_ok.values())
