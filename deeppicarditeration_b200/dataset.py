"""Mirror of the boundary caller `IterableDatasetWithInternalBatch` (picard/dataset.py:20-137).

Only the call contract is kept: the dataset calls `batch_data_generator(n_samples_each_call) -> (tx, y)` and yields
`batch_size`-row views.  The reference's chunking of a buffer into several generator calls exists to dodge the
HBM blow-up of materialised [n*M, 1+nx] path tensors; the fused kernels never materialise them, so a fractional
`n_batch_buffer` is honoured for compatibility but buys nothing here.
"""
from __future__ import annotations

from typing import Callable, Tuple, Union

import torch


class IterableDatasetWithInternalBatch(torch.utils.data.IterableDataset):
    def __init__(self, n: int, n_batch_buffer: Union[int, float], batch_size: int, batch_data_generator: Callable):
        super().__init__()
        (self.n_batch_buffer, self.n_samples_each_buffer, self._n_calls_to_generator_each_buffer,
         self._n_samples_each_call) = self._get_sampling_number_info(n_batch_buffer, batch_size)
        self.batch_size = batch_size
        self.n_buffer_refresh = self._get_n_refresh(n)
        self.batch_data_generator = batch_data_generator
        self.saver = None

    @staticmethod
    def _get_sampling_number_info(n_batch_buffer, batch_size) -> Tuple[int, int, int, int]:  # dataset.py:55-78
        if abs(round(n_batch_buffer) - n_batch_buffer) < 1e-5:
            n_batch_buffer = round(n_batch_buffer)
        if isinstance(n_batch_buffer, int):
            return n_batch_buffer, n_batch_buffer * batch_size, 1, n_batch_buffer * batch_size
        assert isinstance(n_batch_buffer, float) and 0 < n_batch_buffer < 1
        calls = round(1 / n_batch_buffer)
        each = batch_size // calls
        assert each * calls == batch_size, f"batch_size {batch_size} must be a multiple of {calls}"
        return 1, batch_size, calls, each

    def _get_n_refresh(self, n: int) -> int:  # dataset.py:80-99
        k = n // self.n_samples_each_buffer
        assert k * self.n_samples_each_buffer == n, (
            f"Total number of data {n} must be a multiple of batch_size({self.batch_size})*"
            f"n_batch_buffer({self.n_batch_buffer})={self.n_samples_each_buffer}")
        return k

    def set_size(self, n):
        self.n_buffer_refresh = self._get_n_refresh(n)

    def attach_saver(self, saver):
        assert self.saver is None, "Saver can only be attached once"
        self.saver = saver

    def refresh_buffer(self):  # dataset.py:107-116
        xs, ys = [], []
        for _ in range(self._n_calls_to_generator_each_buffer):
            x, y = self.batch_data_generator(self._n_samples_each_call)
            xs.append(x)
            ys.append(y)
        return (xs[0], ys[0]) if len(xs) == 1 else (torch.cat(xs, 0), torch.cat(ys, 0))

    def __iter__(self):  # dataset.py:118-129
        for _ in range(self.n_buffer_refresh):
            buffer = self.refresh_buffer()
            if self.saver is not None:
                self.saver.save(buffer, self.n_samples_each_buffer)
            x, y = buffer
            x = x.view(self.n_batch_buffer, self.batch_size, -1)
            y = y.view(self.n_batch_buffer, self.batch_size, -1)
            for i in range(self.n_batch_buffer):
                yield x[i], y[i]
        if self.saver is not None:
            self.saver.close()

    def __len__(self):
        return self.n_buffer_refresh * self.n_batch_buffer
