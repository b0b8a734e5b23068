"""Orbital bookkeeping for the LCAO hot path: quantum numbers, ground-state occupations and
valence masks per element, and the `ElecInfo` selector that slices them for a model.

The numbers are physical facts (Madelung filling order with the usual exceptions) and have to
agree digit for digit with the reference tables because they index embedding rows
(reference: lcaonet/atomistic/elec.py:6-233 tables, lcaonet/atomistic/info.py:8-127 selector).
They are stored packed: one 18-hex-digit string per element (electrons in each orbital, in
filling order) and one 18-bit mask per element (bit o set = orbital o is a valence orbital).
`tests/test_orbitals.py` checks them against the golden dump of the reference tables.
"""
from __future__ import annotations

import torch
from torch import Tensor

# filling order used by every (.., n_orb) axis in this package
ORBITAL_NAMES = ("1s", "2s", "2p", "3s", "3p", "4s", "3d", "4p", "5s", "4d", "5p", "6s", "4f", "5d", "6p", "7s", "5f", "6d")
_L_OF = {"s": 0, "p": 1, "d": 2, "f": 3}
N_ORB_MAX = len(ORBITAL_NAMES)

# (n, l) per orbital
NL_LIST: Tensor = torch.tensor([[int(nm[0]), _L_OF[nm[1]]] for nm in ORBITAL_NAMES], dtype=torch.long)
# embedding-table height per orbital = capacity 2(2l+1) plus the "0 electrons" row
MAX_ELEC_IDX: Tensor = torch.tensor([2 * (2 * _L_OF[nm[1]] + 1) + 1 for nm in ORBITAL_NAMES], dtype=torch.long)

# ground-state occupations, element Z = row index (row 0 is the dummy element)
_OCCUPATION_HEX = (
    "000000000000000000", "100000000000000000", "200000000000000000", "210000000000000000",
    "220000000000000000", "221000000000000000", "222000000000000000", "223000000000000000",
    "224000000000000000", "225000000000000000", "226000000000000000", "226100000000000000",
    "226200000000000000", "226210000000000000", "226220000000000000", "226230000000000000",
    "226240000000000000", "226250000000000000", "226260000000000000", "226261000000000000",
    "226262000000000000", "226262100000000000", "226262200000000000", "226262300000000000",
    "226261500000000000", "226262500000000000", "226262600000000000", "226262700000000000",
    "226262800000000000", "226261a00000000000", "226262a00000000000", "226262a10000000000",
    "226262a20000000000", "226262a30000000000", "226262a40000000000", "226262a50000000000",
    "226262a60000000000", "226262a61000000000", "226262a62000000000", "226262a62100000000",
    "226262a62200000000", "226262a61400000000", "226262a61500000000", "226262a62500000000",
    "226262a61700000000", "226262a61800000000", "226262a60a00000000", "226262a61a00000000",
    "226262a62a00000000", "226262a62a10000000", "226262a62a20000000", "226262a62a30000000",
    "226262a62a40000000", "226262a62a50000000", "226262a62a60000000", "226262a62a61000000",
    "226262a62a62000000", "226262a62a62010000", "226262a62a62110000", "226262a62a62300000",
    "226262a62a62400000", "226262a62a62500000", "226262a62a62600000", "226262a62a62700000",
    "226262a62a62710000", "226262a62a62900000", "226262a62a62a00000", "226262a62a62b00000",
    "226262a62a62c00000", "226262a62a62d00000", "226262a62a62e00000", "226262a62a62e10000",
    "226262a62a62e20000", "226262a62a62e30000", "226262a62a62e40000", "226262a62a62e50000",
    "226262a62a62e60000", "226262a62a62e70000", "226262a62a61e90000", "226262a62a61ea0000",
    "226262a62a62ea0000", "226262a62a62ea1000", "226262a62a62ea2000", "226262a62a62ea3000",
    "226262a62a62ea4000", "226262a62a62ea5000", "226262a62a62ea6000", "226262a62a62ea6100",
    "226262a62a62ea6200", "226262a62a62ea6201", "226262a62a62ea6202", "226262a62a62ea6221",
    "226262a62a62ea6231", "226262a62a62ea6241", "226262a62a62ea6260", "226262a62a62ea6270",
    "226262a62a62ea6271",
)
# valence-orbital bit masks, element Z = index
_VALENCE_BITS = (
    0x00000, 0x00001, 0x00001, 0x00002, 0x00002, 0x00006, 0x00006, 0x00006, 0x00006, 0x00006, 0x00006, 0x00008,
    0x00008, 0x00018, 0x00018, 0x00018, 0x00018, 0x00018, 0x00018, 0x00020, 0x00020, 0x00060, 0x00060, 0x00060,
    0x00060, 0x00060, 0x00060, 0x00060, 0x00060, 0x00060, 0x00060, 0x000c0, 0x000c0, 0x000c0, 0x000c0, 0x000c0,
    0x000c0, 0x00100, 0x00100, 0x00300, 0x00300, 0x00300, 0x00300, 0x00300, 0x00300, 0x00300, 0x00300, 0x00300,
    0x00300, 0x00600, 0x00600, 0x00600, 0x00600, 0x00600, 0x00600, 0x00800, 0x00800, 0x02800, 0x03800, 0x01800,
    0x01800, 0x01800, 0x01800, 0x01800, 0x03800, 0x01800, 0x01800, 0x01800, 0x01800, 0x01800, 0x01800, 0x03000,
    0x03000, 0x03000, 0x03000, 0x03000, 0x03000, 0x03000, 0x03800, 0x03800, 0x03000, 0x06000, 0x06000, 0x06000,
    0x06000, 0x06000, 0x06000, 0x08000, 0x08000, 0x28000, 0x28000, 0x38000, 0x38000, 0x38000, 0x18000, 0x18000,
    0x38000,
)

ELEC_TABLE: Tensor = torch.tensor([[int(ch, 16) for ch in row] for row in _OCCUPATION_HEX], dtype=torch.long)
VALENCE_TABLE: Tensor = torch.tensor(
    [[(bits >> o) & 1 for o in range(N_ORB_MAX)] for bits in _VALENCE_BITS], dtype=torch.long
)

# highest orbital (index into ORBITAL_NAMES) occupied by any element up to Z, as (Z upper bound, index)
_LAST_ORBITAL_BY_Z = ((2, 0), (4, 1), (10, 2), (12, 3), (18, 4), (20, 5), (30, 6), (36, 7), (38, 8), (48, 9),
                      (54, 10), (56, 11), (80, 13), (86, 14), (88, 15), (96, 17))


class ElecInfo:
    """Which orbitals a model carries and the per-element tables restricted to them.

    Same constructor and properties as the reference selector (info.py:12-127):
    `n_orb = (last orbital index + 1) * n_per_orb`; every table is truncated to the first
    `last+1` orbitals and each orbital column repeated `n_per_orb` times consecutively.
    """

    def __init__(self, max_z: int, max_orb: str | None, min_orb: str | None = None, n_per_orb: int = 1):
        self.max_z, self.max_orb, self.min_orb, self.n_per_orb = max_z, max_orb, min_orb, n_per_orb
        self._min_orb_idx = self._index_of(min_orb) if min_orb else None
        last = self._last_for_z(max_z)
        if max_orb is not None:
            last = max(last, self._index_of(max_orb))
        self._max_orb_idx = last
        self._n_orb = (last + 1) * n_per_orb

    @staticmethod
    def _last_for_z(max_z: int) -> int:
        if max_z <= 0:
            raise ValueError(f"max_z={max_z} is too small.")
        for bound, idx in _LAST_ORBITAL_BY_Z:
            if max_z <= bound:
                return idx
        raise ValueError(f"max_z={max_z} is too large.")

    @staticmethod
    def _index_of(name: str) -> int:
        if name not in ORBITAL_NAMES:
            raise ValueError(f"max_orb={name} is not supported.")
        return ORBITAL_NAMES.index(name)

    def _cols(self, table: Tensor, dim: int) -> Tensor:
        return table.narrow(dim, 0, self._max_orb_idx + 1).repeat_interleave(self.n_per_orb, dim=dim)

    @property
    def n_orb(self) -> int:
        return self._n_orb

    @property
    def min_orb_idx(self) -> int | None:
        if self._min_orb_idx is None:
            return None
        return self._min_orb_idx * self.n_per_orb + (self.n_per_orb - 1)

    @property
    def elec_table(self) -> Tensor:
        return self._cols(ELEC_TABLE[: self.max_z + 1], 1)

    @property
    def valence_table(self) -> Tensor:
        return self._cols(VALENCE_TABLE[: self.max_z + 1], 1)

    @property
    def max_elec_idx(self) -> Tensor:
        return self._cols(MAX_ELEC_IDX, 0)

    @property
    def nl_list(self) -> Tensor:
        return self._cols(NL_LIST, 0)
