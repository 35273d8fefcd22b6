"""Sample sinks (src/io/csv.rs, arrow.rs, parquet.rs).  The CSV writer is host code and is checked here on CPU against
the reference's own expected files and the oracle's restatement; the Arrow / Parquet sinks go through the device
transpose kernel (gpu marker) and are read back with pyarrow."""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def mio():
    import mini_mcmc_b200.io as io

    return io


# ---------------------------------------------------------------- CSV (CPU)
def test_csv_reference_expected_files(mio, tmp_path):
    f = str(tmp_path / "a.csv")
    mio.save_csv(np.zeros((0, 0, 0), dtype=np.float32), f)                      # src/io/csv.rs:160-177
    assert open(f).read().strip() == "chain,observation"
    mio.save_csv(np.array([[[42.0]]]), f)                                        # :180-196
    assert open(f).read().strip() == "chain,observation,dim_0\n0,0,42"
    mio.save_csv(np.array([[[1, 2], [3, 4]], [[10, 20], [30, 40]]]), f)          # :199-217
    assert open(f).read().strip() == "chain,observation,dim_0,dim_1\n0,0,1,2\n0,1,3,4\n1,0,10,20\n1,1,30,40"
    mio.save_csv_tensor(np.array([[[1.0, 2.0], [3.0, 4.0]], [[1.1, 2.1], [3.1, 4.1]]]), f)   # :219-267
    rows = [r.split(",") for r in open(f).read().strip().split("\n")]
    assert rows[0] == ["chain", "observation", "dim_0", "dim_1"]
    assert rows[1:] == [["0", "0", "1", "2"], ["0", "1", "3", "4"], ["1", "0", "1.1", "2.1"], ["1", "1", "3.1", "4.1"]]


@pytest.mark.parametrize("dtype,as_f32", [(np.float64, False), (np.float32, True), (np.uint64, False)])
def test_csv_matches_oracle_restatement(mio, tmp_path, dtype, as_f32):
    rng = np.random.default_rng(3)
    if dtype == np.uint64:
        x = rng.integers(0, 2**40, size=(5, 7, 3)).astype(np.uint64)
    else:
        x = (rng.normal(size=(5, 7, 3)) * 10.0 ** rng.integers(-12, 12, size=(5, 7, 3))).astype(dtype)
        x[0, 0] = [0.0, -0.0, 1.0]
        x[1, 1] = [np.nan, np.inf, -np.inf]
        x[2, 2] = [1e-7, 123456789.0, 0.1]
    f = str(tmp_path / "x.csv")
    (mio.save_csv_tensor if as_f32 else mio.save_csv)(x, f)
    assert open(f).read() == oracle.csv_text(x, as_f32=as_f32)


def test_csv_many_chains_keeps_order(mio, tmp_path):
    x = np.arange(3000 * 4 * 2, dtype=np.float64).reshape(3000, 4, 2)   # several formatting threads
    f = str(tmp_path / "big.csv")
    mio.save_csv(x, f)
    assert open(f).read() == oracle.csv_text(x)


# ---------------------------------------------------------------- Arrow / Parquet (device transpose)
def _check_table(tbl, x, tensor_layout=False):
    import pyarrow as pa

    exp = oracle.long_table(x, tensor_layout=tensor_layout)
    assert tbl.schema.names == list(exp.keys())
    for name in tbl.schema.names:
        field = tbl.schema.field(name)
        assert not field.nullable
        assert field.type == (pa.float64() if name.startswith("dim_") else pa.uint32())
        np.testing.assert_array_equal(tbl.column(name).to_numpy(), exp[name])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((3, 5, 1), np.uint64), ((4, 9, 3), np.float32), ((2, 33, 40), np.float64),
                                         ((70, 2, 100), np.float32), ((5, 4, 33), np.float32)])
def test_arrow_and_parquet_round_trip(mio, cuda_device, tmp_path, shape, dtype):
    import pyarrow as pa
    import pyarrow.parquet as pq
    import torch

    rng = np.random.default_rng(7)
    x = (rng.normal(size=shape) * 100).astype(dtype) if dtype != np.uint64 else rng.integers(0, 1000, size=shape).astype(np.uint64)
    fa, fp, ft = (str(tmp_path / n) for n in ("s.arrow", "s.parquet", "t.parquet"))
    mio.save_arrow(x, fa, rows_per_batch=7 * shape[1])          # several record batches, ragged last one
    with pa.OSFile(fa, "rb") as src:
        _check_table(pa.ipc.open_file(src).read_all(), x)
    dev = torch.from_numpy(x.view(np.int64) if dtype == np.uint64 else x).cuda()
    mio.save_parquet(dev, fp)                                   # CUDA tensor input, one batch
    _check_table(pq.read_table(fp), x)
    mio.save_parquet_tensor(x, ft, rows_per_batch=3 * shape[1])  # [observations, chains, dims] convention
    _check_table(pq.read_table(ft), x, tensor_layout=True)


@pytest.mark.gpu
def test_arrow_empty_sample(mio, cuda_device, tmp_path):
    import pyarrow as pa

    f = str(tmp_path / "e.arrow")
    mio.save_arrow(np.zeros((0, 0, 3), dtype=np.float64), f)    # src/io/arrow.rs: an empty batch with the full schema
    with pa.OSFile(f, "rb") as src:
        tbl = pa.ipc.open_file(src).read_all()
    assert tbl.num_rows == 0 and tbl.schema.names == ["chain", "observation", "dim_0", "dim_1", "dim_2"]
