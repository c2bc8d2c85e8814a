"""Host-side mirror of StVO::GridStructure / GridWindow / getLineCoords.

The reference keeps the bucket grid on the host (stvo-pl/include/gridStructure.h:33-58,
stvo-pl/src/gridStructure.cpp:33-83, stvo-pl/src/lineIterator.cpp:34-77) and so do we: the grid is
built by the caller and handed to the GPU matcher in CSR form.

CSR layout (what include/plmatch.h takes): cell (x, y) with 0 <= x < cols, 0 <= y < rows has the
linear id ``x * rows + y`` -- x outermost, exactly like the reference's ``grid[x][y]``
(gridStructure.cpp:49) -- so one window column is ONE contiguous item range.
``cell_start`` has ``rows * cols + 1`` int32 entries, ``cell_items`` the concatenated buckets in
push order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence, Set, Tuple

import numpy as np

GRID_ROWS = 48  # stvo-pl/include/stereoFrame.h:51
GRID_COLS = 64  # stvo-pl/include/stereoFrame.h:52


@dataclass
class GridWindow:
    """gridStructure.h:35-37 -- (left, right) and (up, down) half-extents in cells."""
    width: Tuple[int, int] = (0, 0)
    height: Tuple[int, int] = (0, 0)

    def as_array(self) -> np.ndarray:
        return np.array([self.width[0], self.width[1], self.height[0], self.height[1]], np.int32)


def getLineCoords(x1: float, y1: float, x2: float, y2: float) -> List[Tuple[int, int]]:
    """gridStructure.cpp:33-41 over LineIterator (lineIterator.cpp:34-77): Bresenham walk from
    double endpoints; start cell / last column come from C-style truncation of the (possibly
    swapped) endpoints, error starts at dx / 2."""
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1 = y1, x1
        x2, y2 = y2, x2
    if x1 > x2:
        x1, x2 = x2, x1
        y1, y2 = y2, y1
    dx = x2 - x1
    dy = abs(y2 - y1)
    error = dx / 2.0
    ystep = 1 if y1 < y2 else -1
    x, y, max_x = int(x1), int(y1), int(x2)
    out = []
    while x <= max_x:
        out.append((y, x) if steep else (x, y))
        error -= dy
        if error < 0:
            y += ystep
            error += dx
        x += 1
    return out


class GridStructure:
    """gridStructure.h:39-58.  ``at`` returns a sink list for off-grid coordinates
    (gridStructure.cpp:56-63); ``get`` clamps the window to the grid (:65-76)."""

    def __init__(self, rows: int, cols: int):
        if rows <= 0 or cols <= 0:
            raise RuntimeError("[GridStructure] invalid dimension")
        self.rows, self.cols = int(rows), int(cols)
        self._grid: List[List[List[int]]] = [[[] for _ in range(self.rows)] for _ in range(self.cols)]
        self._out_of_bounds: List[int] = []

    def at(self, x, y) -> List[int]:
        x, y = int(x), int(y)  # C++ double -> int conversion truncates toward zero
        if 0 <= x < self.cols and 0 <= y < self.rows:
            return self._grid[x][y]
        return self._out_of_bounds

    def get(self, x: int, y: int, w: GridWindow, indices: Set[int]) -> None:
        min_x = max(0, x - w.width[0])
        max_x = min(self.cols, x + w.width[1] + 1)
        min_y = max(0, y - w.height[0])
        max_y = min(self.rows, y + w.height[1] + 1)
        for x_ in range(min_x, max_x):
            for y_ in range(min_y, max_y):
                indices.update(self._grid[x_][y_])

    def clear(self) -> None:
        for col in self._grid:
            for cell in col:
                cell.clear()

    def to_csr(self) -> Tuple[np.ndarray, np.ndarray]:
        counts = np.fromiter((len(self._grid[x][y]) for x in range(self.cols) for y in range(self.rows)),
                             dtype=np.int64, count=self.rows * self.cols)
        cell_start = np.zeros(self.rows * self.cols + 1, np.int32)
        np.cumsum(counts, out=cell_start[1:])
        items = [i for x in range(self.cols) for y in range(self.rows) for i in self._grid[x][y]]
        return cell_start, np.asarray(items, np.int32)


def csr_from_cells(cell_x: np.ndarray, cell_y: np.ndarray, item: np.ndarray, rows: int, cols: int):
    """CSR for an arbitrary (cell, item) push sequence: items land in their cell in push order,
    off-grid pushes go to the reference's out_of_bounds sink, i.e. are dropped."""
    cell_x = np.asarray(cell_x, np.int64)
    cell_y = np.asarray(cell_y, np.int64)
    item = np.asarray(item, np.int32)
    ok = (cell_x >= 0) & (cell_x < cols) & (cell_y >= 0) & (cell_y < rows)
    cid = cell_x[ok] * rows + cell_y[ok]
    order = np.argsort(cid, kind="stable")
    counts = np.bincount(cid, minlength=rows * cols)
    cell_start = np.zeros(rows * cols + 1, np.int32)
    np.cumsum(counts, out=cell_start[1:])
    return cell_start, item[ok][order]


def csr_from_points(px: np.ndarray, py: np.ndarray, rows: int = GRID_ROWS, cols: int = GRID_COLS):
    """``grid.at(px, py).push_back(idx)`` for idx = 0..n-1 (stereoFrame.cpp:146-150): px/py are the
    already-scaled double coordinates; truncation toward zero picks the cell."""
    cx = np.trunc(np.asarray(px, np.float64)).astype(np.int64)
    cy = np.trunc(np.asarray(py, np.float64)).astype(np.int64)
    return csr_from_cells(cx, cy, np.arange(len(cx), dtype=np.int32), rows, cols)


def line_cells(x1, y1, x2, y2):
    """Vectorised getLineCoords over many segments.  Returns (line_id, cell_x, cell_y) in the
    reference's push order (line by line, Bresenham order inside a line)."""
    x1 = np.array(x1, np.float64); y1 = np.array(y1, np.float64)
    x2 = np.array(x2, np.float64); y2 = np.array(y2, np.float64)
    n = len(x1)
    steep = np.abs(y2 - y1) > np.abs(x2 - x1)
    x1, y1 = np.where(steep, y1, x1), np.where(steep, x1, y1)
    x2, y2 = np.where(steep, y2, x2), np.where(steep, x2, y2)
    swap = x1 > x2
    x1, x2 = np.where(swap, x2, x1), np.where(swap, x1, x2)
    y1, y2 = np.where(swap, y2, y1), np.where(swap, y1, y2)
    dx = x2 - x1
    dy = np.abs(y2 - y1)
    error = dx / 2.0
    ystep = np.where(y1 < y2, 1, -1).astype(np.int64)
    x = np.trunc(x1).astype(np.int64)
    y = np.trunc(y1).astype(np.int64)
    max_x = np.trunc(x2).astype(np.int64)
    ids, cxs, cys, steps = [], [], [], []
    alive = np.arange(n)
    step = 0
    while True:
        live = x[alive] <= max_x[alive]
        alive = alive[live]
        if alive.size == 0:
            break
        st = steep[alive]
        cxs.append(np.where(st, y[alive], x[alive]))
        cys.append(np.where(st, x[alive], y[alive]))
        ids.append(alive.copy())
        steps.append(np.full(alive.size, step, np.int64))
        error[alive] -= dy[alive]
        neg = error[alive] < 0
        idx_neg = alive[neg]
        y[idx_neg] += ystep[idx_neg]
        error[idx_neg] += dx[idx_neg]
        x[alive] += 1
        step += 1
    if not ids:
        z = np.zeros(0, np.int64)
        return z, z, z
    ids = np.concatenate(ids); cxs = np.concatenate(cxs); cys = np.concatenate(cys); steps = np.concatenate(steps)
    order = np.lexsort((steps, ids))
    return ids[order], cxs[order], cys[order]


def csr_from_lines(x1, y1, x2, y2, rows: int = GRID_ROWS, cols: int = GRID_COLS):
    """Grid fill of stereoFrame.cpp:336-349: every line index is pushed into each cell of its
    Bresenham walk (scaled double endpoints)."""
    ids, cx, cy = line_cells(x1, y1, x2, y2)
    return csr_from_cells(cx, cy, ids.astype(np.int32), rows, cols)


def line_directions(x1, y1, x2, y2) -> np.ndarray:
    """directions2 as the callers build it (stereoFrame.cpp:342-344): normalize((ex-sx, ey-sy)) with
    matching.h:43-48 (divide by sqrt(x*x + y*y); zero-length lines give NaN, which the matcher's
    direction filter lets through)."""
    vx = np.asarray(x2, np.float64) - np.asarray(x1, np.float64)
    vy = np.asarray(y2, np.float64) - np.asarray(y1, np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mag = np.sqrt(vx * vx + vy * vy)
        return np.stack([vx / mag, vy / mag], axis=1)
