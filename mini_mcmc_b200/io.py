"""`mini_mcmc::io` surface: save_csv, save_csv_tensor, save_arrow, save_parquet, save_parquet_tensor
(src/io/csv.rs:47-147, src/io/arrow.rs:53-117, src/io/parquet.rs:49-221).

Long format, one row per (chain, observation): `chain: u32, observation: u32, dim_0..dim_{D-1}: f64` (non-nullable).
Samples may be host arrays or CUDA tensors.  For Arrow / Parquet the f64 columns are produced on the device
(csrc/mmc_sink.cu), copied block by block into pinned staging and handed to pyarrow as zero-copy buffers, so a
sample larger than host memory budget streams through two staging buffers; the file encoders are pyarrow (library
code, as the arrow / parquet crates are in the reference).  CSV is written natively (csrc/mmc_sink_csv.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_NP_DT = {np.dtype(np.float32): L.MMC_F32, np.dtype(np.float64): L.MMC_F64, np.dtype(np.uint64): L.MMC_U64}


def _host_array(data):
    if hasattr(data, "data_ptr"):   # torch tensor
        import torch

        if data.dtype == torch.int64:
            return data.detach().cpu().contiguous().numpy().view(np.uint64)
        data = data.detach().cpu().contiguous().numpy()
    a = np.asarray(data)
    if a.ndim != 3:
        raise ValueError(f"expected [chains, observations, dims], got shape {a.shape}")
    if a.dtype in (np.dtype(np.int64), np.dtype(np.int32), np.dtype(np.uint32)):
        if a.size and a.min() < 0:
            raise ValueError("negative integer states are not supported by the CSV sink")
        a = a.astype(np.uint64)
    if a.dtype not in _NP_DT:
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def save_csv(data, filename: str) -> None:
    """save_csv(&Array3<T>, filename), src/io/csv.rs:47-77."""
    a = _host_array(data)
    c, n, d = a.shape
    L.check(L.lib.mmc_save_csv(L.vp(a) if a.size else None, C.c_int32(_NP_DT[a.dtype]), C.c_int64(c), C.c_int64(n), C.c_int32(d),
                               filename.encode()))


def save_csv_tensor(tensor, filename: str) -> None:
    """save_csv_tensor(Tensor<B, 3>, filename), src/io/csv.rs:110-147: values are formatted as f32."""
    if hasattr(tensor, "data_ptr"):
        import torch

        tensor = tensor.detach().to(torch.float32).cpu().numpy()
    save_csv(np.asarray(tensor, dtype=np.float32), filename)


def _schema(n_dims: int, first=("chain", "observation")):
    import pyarrow as pa

    fields = [pa.field(first[0], pa.uint32(), nullable=False), pa.field(first[1], pa.uint32(), nullable=False)]
    fields += [pa.field(f"dim_{i}", pa.float64(), nullable=False) for i in range(n_dims)]
    return pa.schema(fields)


def _to_device(data):
    import torch

    if hasattr(data, "data_ptr"):
        t = data
    else:
        a = np.asarray(data)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        elif a.dtype not in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int64)):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dim() != 3:
        raise ValueError(f"expected a 3-d sample, got shape {tuple(t.shape)}")
    if t.dtype not in (torch.float32, torch.float64, torch.int64):
        t = t.to(torch.float64)
    return t.to("cuda").contiguous()


def record_batches(data, rows_per_batch: int = 1 << 22, tensor_layout: bool = False, zero_copy: bool = False):
    """Yields pyarrow RecordBatches of the long-format table.  tensor_layout=True is save_parquet_tensor's
    convention: data is [observations, chains, dims] and the columns start with observation, chain.
    zero_copy=True wraps the two reused pinned staging buffers directly: a batch is then only valid until the generator
    is resumed twice (the writers below consume each batch before asking for the next); the default copies the columns,
    so batches may be collected (list(record_batches(x)), pa.Table.from_batches)."""
    import pyarrow as pa
    import torch

    x = _to_device(data)
    outer, inner, d = x.shape   # rows are (outer, inner) pairs in row-major order
    names = ("observation", "chain") if tensor_layout else ("chain", "observation")
    schema = _schema(d, names)
    if outer * inner == 0 or d == 0:
        cols = [pa.array(np.empty(0, dtype=np.uint32)), pa.array(np.empty(0, dtype=np.uint32))] + \
               [pa.array(np.empty(0, dtype=np.float64)) for _ in range(d)]
        if outer * inner and d == 0:
            idx_o = np.repeat(np.arange(outer, dtype=np.uint32), inner)
            idx_i = np.tile(np.arange(inner, dtype=np.uint32), outer)
            cols = [pa.array(idx_o), pa.array(idx_i)]
        yield pa.RecordBatch.from_arrays(cols, schema=schema)
        return
    dt = {torch.float32: L.MMC_F32, torch.float64: L.MMC_F64, torch.int64: L.MMC_U64}[x.dtype]
    per = max(1, min(outer, rows_per_batch // max(inner, 1)))
    rows_max = per * inner
    dev = [torch.empty(d * rows_max, dtype=torch.float64, device="cuda") for _ in range(2)]
    pin = [torch.empty(d * rows_max, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    stream = torch.cuda.current_stream()
    blocks = [(o0, min(per, outer - o0)) for o0 in range(0, outer, per)]

    def launch(i):
        o0, cnt = blocks[i]
        b = i & 1
        L.check(L.lib.mmc_sink_columns_dev(L.vp(x), C.c_int32(dt), C.c_int64(outer), C.c_int64(inner), C.c_int32(d), C.c_int64(o0),
                                           C.c_int64(cnt), L.vp(dev[b]), C.c_void_p(stream.cuda_stream)))
        pin[b][: d * cnt * inner].copy_(dev[b][: d * cnt * inner], non_blocking=True)
        ev[b].record(stream)

    launch(0)
    for i, (o0, cnt) in enumerate(blocks):
        if i + 1 < len(blocks):
            launch(i + 1)      # transposes + copies the next block while this one is encoded
        ev[i & 1].synchronize()
        rows = cnt * inner
        host = pin[i & 1].numpy()
        idx_o = np.repeat(np.arange(o0, o0 + cnt, dtype=np.uint32), inner)
        idx_i = np.tile(np.arange(inner, dtype=np.uint32), cnt)
        cols = [pa.array(idx_o), pa.array(idx_i)] + \
               [pa.array(host[k * rows:(k + 1) * rows] if zero_copy else host[k * rows:(k + 1) * rows].copy()) for k in range(d)]
        yield pa.RecordBatch.from_arrays(cols, schema=schema)


def save_arrow(data, filename: str, rows_per_batch: int = 1 << 22) -> None:
    """save_arrow(&Array3<T>, filename), src/io/arrow.rs:53-117: Arrow IPC file."""
    import pyarrow as pa

    it = record_batches(data, rows_per_batch, zero_copy=True)
    first = next(it)
    with pa.OSFile(filename, "wb") as sink, pa.ipc.new_file(sink, first.schema) as w:
        w.write_batch(first)
        for b in it:
            w.write_batch(b)


def _save_parquet(data, filename, rows_per_batch, tensor_layout):
    import pyarrow.parquet as pq

    it = record_batches(data, rows_per_batch, tensor_layout=tensor_layout, zero_copy=True)
    first = next(it)
    with pq.ParquetWriter(filename, first.schema) as w:
        w.write_batch(first)
        for b in it:
            w.write_batch(b)


def save_parquet(data, filename: str, rows_per_batch: int = 1 << 22) -> None:
    """save_parquet(&Array3<T>, filename), src/io/parquet.rs:49-122."""
    _save_parquet(data, filename, rows_per_batch, False)


def save_parquet_tensor(tensor, filename: str, rows_per_batch: int = 1 << 22) -> None:
    """save_parquet_tensor(&Tensor<B, 3>, filename), src/io/parquet.rs:154-221: tensor is
    [observations, chains, dims]; columns observation, chain, dim_*."""
    _save_parquet(tensor, filename, rows_per_batch, True)
