"""Stand-in for the slice of torch_sparse.SparseTensor the reference touches
(lcaonet.py:462-477): ctor (sort by row*ncols+col only when unsorted, default non-stable
argsort), __getitem__(LongTensor) = CSR row select, set_value(None).sum(dim=1) = row counts,
storage.row()/col()/value().  Test infrastructure."""
import torch


class _Storage:
    def __init__(self, row, col, value):
        self._row, self._col, self._value = row, col, value

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value


class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None, _sorted=False):
        n_rows, n_cols = sparse_sizes
        if not _sorted:
            key = row * n_cols + col
            if key.numel() > 1 and bool((key[1:] < key[:-1]).any()):
                perm = key.argsort()
                row, col = row[perm], col[perm]
                value = value[perm] if value is not None else None
        self.storage = _Storage(row, col, value)
        self._sizes = (n_rows, n_cols)

    def _rowptr(self):
        cnt = torch.zeros(self._sizes[0], dtype=torch.long, device=self.storage._row.device)
        cnt.scatter_add_(0, self.storage._row, torch.ones_like(self.storage._row))
        ptr = torch.zeros(self._sizes[0] + 1, dtype=torch.long, device=cnt.device)
        ptr[1:] = cnt.cumsum(0)
        return ptr, cnt

    def __getitem__(self, idx):
        ptr, cnt = self._rowptr()
        c = cnt[idx]
        new_row = torch.arange(idx.numel(), device=idx.device).repeat_interleave(c)
        start = ptr[idx].repeat_interleave(c)
        first = (c.cumsum(0) - c).repeat_interleave(c)
        pos = start + (torch.arange(new_row.numel(), device=idx.device) - first)
        val = self.storage._value[pos] if self.storage._value is not None else None
        return SparseTensor(new_row, self.storage._col[pos], val, (idx.numel(), self._sizes[1]), _sorted=True)

    def set_value(self, value, layout=None):
        return SparseTensor(self.storage._row, self.storage._col, value, self._sizes, _sorted=True)

    def sum(self, dim):
        assert dim == 1
        if self.storage._value is None:
            return self._rowptr()[1].to(torch.float)
        out = torch.zeros(self._sizes[0], dtype=self.storage._value.dtype, device=self.storage._row.device)
        return out.scatter_add_(0, self.storage._row, self.storage._value)
