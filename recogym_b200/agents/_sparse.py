"""Host-side sparse form of the count agents' P x P statistics (num_products too large for dense matrices: BASELINE
configs[3], P = 100 000 -> 3 x 80 GB).  Same update rules and the same numpy.argmax results as the dense arrays of
the reference (recogym/agents/organic_count.py:49-82, recogym/agents/bandit_count.py:22-62); the device counterpart is
`recogym_b200.logs.SparseCountStats` (hash tables filled by scatter kernels)."""
import numpy as np


class HostSparseCounts:
    """rows[i][j] = [a, b] (co_counts | pulls, clicks); row_a / row_b = increments that apply to EVERY cell of a row
    (numpy's `pulls_a[None, jx] += 1` while BanditCount.last_product_viewed is None)."""

    def __init__(self, P):
        self.P = int(P)
        self.rows = {}
        self.row_a = {}
        self.row_b = {}

    def add(self, i, j, da, db=0):
        cell = self.rows.setdefault(int(i), {}).setdefault(int(j), [0, 0])
        cell[0] += da
        cell[1] += db

    def add_row(self, i, da, db=0):
        self.row_a[int(i)] = self.row_a.get(int(i), 0) + da
        self.row_b[int(i)] = self.row_b.get(int(i), 0) + db

    def n_cells(self):
        return sum(len(r) for r in self.rows.values())

    # ---- numpy.argmax over the dense row (first maximum wins) ------------------------------------------------------
    def organic_argmax(self, i):
        """co_counts[i, :].argmax(): the largest count, lowest column among ties; an all-zero row -> 0."""
        best_n, best_j = 0, 0
        for j, (n, _) in self.rows.get(int(i), {}).items():
            if n > best_n or (n == best_n and n > 0 and j < best_j):
                best_n, best_j = n, j
        return best_j

    def bandit_argmax(self, i):
        """((clicks + 1) / (pulls + 2))[i, :].argmax() and that maximum.  Untouched cells hold the row's base value
        (row_b + 1) / (row_a + 2): the best touched cell wins when it beats the base or when the whole row has been
        touched; otherwise the lowest column whose cell is not below the base."""
        i = int(i)
        rp, rc = self.row_a.get(i, 0), self.row_b.get(i, 0)
        base = (rc + 1) / (rp + 2)
        row = self.rows.get(i)
        if not row:
            return 0, base
        val = {j: (b + rc + 1) / (a + rp + 2) for j, (a, b) in row.items()}
        vmax = max(val.values())
        if vmax > base or len(val) >= self.P:
            return min(j for j, x in val.items() if x == vmax), vmax
        j = 0
        while j in val and val[j] < base:
            j += 1
        return j, base

    def organic_table(self):
        table = np.zeros(self.P, np.int32)
        for i in self.rows:
            table[i] = self.organic_argmax(i)
        return table

    def bandit_table(self):
        table = np.zeros(self.P, np.int32)
        ps = np.full(self.P, 0.5)
        for i in set(self.rows) | set(self.row_a):
            table[i], ps[i] = self.bandit_argmax(i)
        return table, ps

    # ---- device residency -------------------------------------------------------------------------------------------
    def arrays(self):
        """-> keys (uint64 row << 32 | col), a, b as numpy arrays (any order)."""
        n = self.n_cells()
        keys, a, b = np.empty(n, np.uint64), np.empty(n, np.int64), np.empty(n, np.int64)
        k = 0
        for i, row in self.rows.items():
            for j, (x, y) in row.items():
                keys[k] = (np.uint64(i) << np.uint64(32)) | np.uint64(j)
                a[k], b[k] = x, y
                k += 1
        return keys, a, b

    @classmethod
    def from_cells(cls, P, cells, row_a=None, row_b=None):
        """cells: {(row, col): (a, b)} as _DeviceHash.cells() returns them; row_a / row_b: dense [P] arrays or None."""
        s = cls(P)
        for (i, j), (x, y) in cells.items():
            s.rows.setdefault(i, {})[j] = [x, y]
        if row_a is not None:
            for i in np.flatnonzero(np.asarray(row_a) | np.asarray(row_b)):
                s.row_a[int(i)] = int(row_a[i])
                s.row_b[int(i)] = int(row_b[i])
        return s
