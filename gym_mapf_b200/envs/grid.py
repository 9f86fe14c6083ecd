"""Obstacle grid (reference gym_mapf/envs/grid.py).  Same indexing, iteration order and errors; the cells are
kept as one uint8 mask so the engine can take the grid without re-parsing it."""
import numpy as np


class ObstacleCell:
    pass


class EmptyCell:
    pass


CHAR_TO_CELL = {".": EmptyCell, "@": ObstacleCell}
_KINDS = (EmptyCell, ObstacleCell)


class MapfGrid:
    def __init__(self, map_lines):
        rows = []
        for line in map_lines:
            # an unknown character raises KeyError, as in the reference (grid.py:21)
            rows.append([0 if CHAR_TO_CELL[ch] is EmptyCell else 1 for ch in line.strip()])
        self.obstacles = np.array(rows, dtype=np.uint8).reshape(len(rows), -1)
        self.max_row = len(rows) - 1
        self.max_col = len(rows[0]) - 1

    def _row(self, r):
        if not -len(self) <= r < len(self):
            raise IndexError("list index out of range")
        return [_KINDS[v] for v in self.obstacles[r]]

    def __getitem__(self, key):
        # grid[r] -> the row (a list of cell classes); grid[r, c] / grid[(r, c)] -> one cell class (grid.py:27-35)
        if type(key) == int:
            return self._row(key)
        if len(key) == 2:
            r, c = key
            h, w = self.obstacles.shape
            if not (-h <= r < h and -w <= c < w):
                raise IndexError("list index out of range")
            return _KINDS[self.obstacles[r, c]]
        out = self._row(key[0]) if len(key) else self
        for idx in key[1:]:
            out = out[idx]
        return out

    def __iter__(self):
        # COLUMN-major: this order defines the cell numbering of the env (grid.py:37-40, mapf_env.py:142-143)
        for c in range(self.obstacles.shape[1]):
            for r in range(self.obstacles.shape[0]):
                yield (r, c)

    def __len__(self):
        return self.obstacles.shape[0]

    def __eq__(self, other):
        return self.obstacles.shape == other.obstacles.shape and bool((self.obstacles == other.obstacles).all())

    def free_cells(self):
        """Free cells in column-major order as an int array [L, 2] of (row, col)."""
        cols, rows = np.nonzero(self.obstacles.T == 0)
        return np.stack([rows, cols], axis=1)
