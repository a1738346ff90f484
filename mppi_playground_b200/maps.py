"""Host half of the device map rasteriser (SURVEY 8f "next" row 4, C ABI ``mppi_raster_map``).

The reference builds its occupancy grids with Python loops over cells
(``ObstacleMap.add_circle_obstacle`` src/envs/obstacle_map_2d.py:103-123, ``add_rectangle_obstacle`` :125-160) and
a Euclidean distance transform (``LaneMap.populate_map`` src/envs/lane_map_2d.py:68-88): seconds at 800 x 800, so
maps cannot follow moving obstacles. Here the *shape -> cell* conversions stay on the host exactly as the
reference writes them (fp64 numpy, a handful of scalars per shape) and the painting runs on the device, straight
into the packed layout the solve kernels read (``MPPI.rasterise_map``). There is no CPU painter in this package:
the numpy restatement used by the tests lives in ``oracle/``.
"""
from __future__ import annotations

from math import ceil
from typing import List, Sequence, Tuple

import numpy as np


class ObstacleRaster:
    """Shape list of one ``ObstacleMap``; same constructor arguments and ``add_*`` calls as the reference class."""

    mode = 0  # _capi.RASTER_OBSTACLE

    def __init__(self, map_size: Tuple[int, int] = (20, 20), cell_size: float = 0.01) -> None:
        assert len(map_size) == 2 and cell_size > 0  # obstacle_map_2d.py:70-73
        assert map_size[0] % 2 == 0 and map_size[1] % 2 == 0
        self.width, self.height = ceil(map_size[0] / cell_size), ceil(map_size[1] / cell_size)  # :75-77
        self.cell_size = cell_size
        self.origin = np.array([self.width / 2, self.height / 2]).astype(int)  # :83-86
        self.x_lim = [-cell_size * self.width / 2, cell_size * self.width / 2]  # :92-96
        self.y_lim = [-cell_size * self.height / 2, cell_size * self.height / 2]
        self.discs: List[Tuple[int, int, int]] = []
        self.rects: List[Tuple[int, int, int, int]] = []

    def add_circle_obstacle(self, center: Sequence[float], radius: float) -> None:
        assert len(center) == 2 and radius > 0
        c = np.round((np.asarray(center, dtype=np.float64) / self.cell_size) + self.origin).astype(int)  # :112-114
        r = ceil(radius / self.cell_size)  # :115
        self.discs.append((int(c[0]), int(c[1]), r * r))  # painted where i**2 + j**2 <= radius_occ**2 (:120)

    def add_rectangle_obstacle(self, center: Sequence[float], width: float, height: float) -> None:
        assert len(center) == 2 and width > 0 and height > 0
        c = np.ceil((np.asarray(center, dtype=np.float64) / self.cell_size) + self.origin).astype(int)  # :139-141
        w_occ, h_occ = ceil(width / self.cell_size), ceil(height / self.cell_size)
        x0, x1 = c[0] - ceil(w_occ / 2), c[0] + ceil(w_occ / 2)  # :145-148
        y0, y1 = c[1] - ceil(h_occ / 2), c[1] + ceil(h_occ / 2)
        x0, x1 = (int(np.clip(v, 0, self.width - 1)) for v in (x0, x1))  # :151-154
        y0, y1 = (int(np.clip(v, 0, self.height - 1)) for v in (y0, y1))
        self.rects.append((x0, x1, y0, y1))  # map[x0:x1, y0:y1] = 1 (:156)

    def clear(self) -> None:
        self.discs, self.rects = [], []

    @classmethod
    def from_obstacle_map(cls, om, map_size: Tuple[int, int]) -> "ObstacleRaster":
        """From a reference ``ObstacleMap`` (its ``circle_obs_list`` / ``rectangle_obs_list`` keep every shape)."""
        r = cls(map_size=map_size, cell_size=om._cell_size)
        for c in om.circle_obs_list:
            r.add_circle_obstacle(c.center, c.radius)
        for q in om.rectangle_obs_list:
            r.add_rectangle_obstacle(q.center, q.width, q.height)
        return r


class LaneRaster:
    """Centre-line cells and the squared cell radius of one ``LaneMap`` (lane_map_2d.py:51-88)."""

    mode = 1  # _capi.RASTER_LANE

    def __init__(self, lane: np.ndarray, lane_width: float, map_size: Tuple[int, int] = (20, 20),
                 cell_size: float = 0.01) -> None:
        lane = np.asarray(lane)
        assert lane_width > 0 and lane.ndim == 2 and lane.shape[1] == 3  # :33-34
        self.width, self.height = ceil(map_size[0] / cell_size), ceil(map_size[1] / cell_size)  # :55
        self.cell_size = cell_size
        self.origin = np.array([self.width // 2, self.height // 2])  # :58
        self.x_lim = [-map_size[0] / 2, map_size[0] / 2]  # :65-66
        self.y_lim = [-map_size[1] / 2, map_size[1] / 2]
        # :73-78 - int(round(x / cell)) + origin, kept when inside the map
        cx = np.rint(lane[:, 0].astype(np.float64) / cell_size).astype(np.int64) + int(self.origin[0])
        cy = np.rint(lane[:, 1].astype(np.float64) / cell_size).astype(np.int64) + int(self.origin[1])
        keep = (cx >= 0) & (cx < self.width) & (cy >= 0) & (cy < self.height)
        cells = np.unique(np.stack([cx[keep], cy[keep]], axis=1), axis=0)
        # :81-83 - distance_transform_edt(map) <= (lane_width / 2) / cell: the EDT of a cell is sqrt(d2) of an
        # integer squared cell distance in fp64, so the test is d2 <= r2 for the largest integer r2 it accepts
        max_distance = (lane_width / 2) / cell_size
        r2 = int(max_distance * max_distance)
        while np.sqrt(np.float64(r2 + 1)) <= max_distance:
            r2 += 1
        while r2 >= 0 and np.sqrt(np.float64(r2)) > max_distance:
            r2 -= 1
        self.r2 = r2
        self.discs = [(int(x), int(y), r2) for x, y in cells] if r2 >= 0 else []
        self.rects: List[Tuple[int, int, int, int]] = []


def flatten_shapes(raster) -> Tuple[np.ndarray, np.ndarray]:
    """(discs [n,3] int32, rects [m,4] int32) in the layout ``mppi_raster_map`` takes."""
    discs = np.asarray(raster.discs, dtype=np.int32).reshape(-1, 3)
    rects = np.asarray(raster.rects, dtype=np.int32).reshape(-1, 4)
    return np.ascontiguousarray(discs), np.ascontiguousarray(rects)
