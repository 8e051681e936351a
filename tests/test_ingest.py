"""Cutout ingest: the native parallel chunk decoder (zlib + HDF5 byte
shuffle, csrc/decode.cu) and the lazily loaded cutouts built on it.  The chunk ENCODING is
the one the reference writes (data.py:245-248); the encoder used here is Python's own zlib +
a NumPy transpose, i.e. independent of the decoder under test."""

import json
import os
import zlib

import numpy as np
import pytest

import atlite_b200 as ab
from atlite_b200 import _lib, ingest, synthetic as syn


@pytest.fixture(scope="module")
def wind_arrays():
    ds = syn.make_dataset(70, 45, 230, kinds=("wind",))
    return ds, {k: np.asarray(ds.raw(k)) for k in ds.keys()}


@pytest.mark.parametrize("chunk,level,shuffle", [((100, None, None), 9, True), ((64, 32, 40), 6, True),
                                                  ((7, 45, 70), 1, False), ((50, 16, 16), 0, True)])
def test_chunk_decoder_round_trips_any_time_window(tmp_path, wind_arrays, chunk, level, shuffle):
    ds, arrs = wind_arrays
    p = ingest.write_chunked(tmp_path / "c.bin", arrs, dict(ds.coords), chunk=chunk, complevel=level, shuffle=shuffle)
    lz = ingest.open_cutout(p)
    assert lz.lazy and list(lz.coords["time"]) == list(ds.coords["time"])
    for lo, hi in ((0, 230), (37, 212), (99, 101), (229, 230), (5, 5)):
        sub = lz.isel_time(lo, hi)
        for k, a in arrs.items():
            got = sub.raw(k)
            assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
            np.testing.assert_array_equal(got, a[lo:hi])


def test_shuffle_filter_matches_the_hdf5_definition():
    """H5Zshuffle: byte k of element i goes to position k * n + i."""
    a = np.array([0x01020304, 0x0A0B0C0D], dtype="<u4")
    assert ingest.shuffle_bytes(a) == bytes([0x04, 0x0D, 0x03, 0x0C, 0x02, 0x0B, 0x01, 0x0A])


def test_packed_big_endian_integers_and_fill_values(tmp_path, wind_arrays):
    ds, arrs = wind_arrays
    v = np.round(arrs["wnd100m"] * 100).astype(">i2")
    v[3, 4, 5] = -32767
    p = ingest.write_chunked(tmp_path / "d.bin", {"w": v}, dict(ds.coords), chunk=(64, 45, 70))
    idx = json.load(open(str(p) + ".index.json"))["variables"]["w"]
    cv = ingest.ChunkedVariable(p, **{k: idx[k] for k in ("shape", "chunk", "dtype", "offsets", "sizes", "origins",
                                                          "shuffle", "deflate")},
                                scale_factor=0.01, add_offset=1.0, fill_value=-32767)
    got = cv.load(0, 100)
    want = v[:100].astype(np.float32) * np.float32(0.01) + np.float32(1.0)
    want[3, 4, 5] = np.nan
    np.testing.assert_array_equal(got, want)


def test_corrupt_chunks_are_reported(tmp_path, wind_arrays):
    ds, arrs = wind_arrays
    p = ingest.write_chunked(tmp_path / "e.bin", {"wnd100m": arrs["wnd100m"]}, dict(ds.coords), chunk=(100, None, None))
    raw = bytearray(open(p, "rb").read())
    raw[100:140] = b"\x00" * 40
    open(p, "wb").write(raw)
    with pytest.raises(_lib.AtlError, match="zlib"):
        ingest.open_cutout(p).isel_time(0, 10)
    with pytest.raises(_lib.AtlError, match="cannot open"):
        ingest.ChunkedVariable(tmp_path / "missing.bin", (10, 45, 70), (10, 45, 70), "<f4", [0], [10], [[0, 0, 0]]).load(0, 5)


def test_decoder_is_the_library_not_python(tmp_path, wind_arrays):
    """One chunk, decoded through the C ABI directly."""
    import ctypes as C

    a = np.arange(2 * 3 * 4, dtype="<f4").reshape(2, 3, 4)
    comp = zlib.compress(ingest.shuffle_bytes(a), 9)
    f = tmp_path / "one.bin"
    f.write_bytes(b"junk" + comp)
    spec = _lib.ChunkSpec()
    spec.ny, spec.nx, spec.elem_bytes, spec.shuffle, spec.deflate = 3, 4, 4, 1, 1
    spec.chunk[0], spec.chunk[1], spec.chunk[2] = 2, 3, 4
    out = np.full((1, 3, 4), -1, dtype=np.float32)
    offs, sizes, orig = np.array([4], np.int64), np.array([len(comp)], np.int64), np.zeros((1, 3), np.int64)
    _lib.check(_lib.load().atl_decode_chunks(str(f).encode(), C.byref(spec), 1, _lib.ptr(offs), _lib.ptr(sizes),
                                             _lib.ptr(orig), 1, 1, out.ctypes.data_as(C.c_void_p), 2))
    np.testing.assert_array_equal(out[0], a[1])


def test_netcdf3_classic_files_open_lazily(tmp_path, wind_arrays):
    from scipy.io import netcdf_file

    ds, arrs = wind_arrays
    p = str(tmp_path / "c3.nc")
    f = netcdf_file(p, "w")
    for n, k in (("time", 230), ("y", 45), ("x", 70)):
        f.createDimension(n, k)
    for n in ("x", "y"):
        f.createVariable(n, "d", (n,))[:] = np.asarray(ds.coords[n])
    t = f.createVariable("time", "i", ("time",))
    t[:] = np.arange(230)
    t.units = "hours since 2013-01-01 00:00:00"
    for n, a in arrs.items():
        f.createVariable(n, "f", ("time", "y", "x"))[:] = a
    f.close()
    c = ab.Cutout(path=p) if not ab.labelled.HAVE_XARRAY else ab.Cutout(data=ingest.open_netcdf3(p))
    assert c.data.lazy and c.shape == (45, 70)
    np.testing.assert_array_equal(c.data.isel_time(100, 130).raw("roughness"), arrs["roughness"][100:130])
    assert list(c.coords["time"][:2]) == list(ds.coords["time"][:2])


@pytest.mark.gpu
def test_cutout_from_a_chunk_file_converts_like_the_in_memory_cutout(tmp_path, wind_arrays, monkeypatch):
    from atlite_b200 import convert as cv

    ds, arrs = wind_arrays
    p = ingest.write_chunked(tmp_path / "cut.bin", arrs, dict(ds.coords), chunk=(50, None, None))
    m = syn.make_shapes(70, 45, 11)
    lazy = ab.Cutout(data=ingest.open_cutout(p, pinned=True))
    monkeypatch.setattr(cv, "PART_BYTES", 100 * cv._bytes_per_step(lazy.data))
    a = ab.Cutout(data=ds).wind("Vestas_V112_3MW", matrix=m, aggregate_time=None)
    b = lazy.wind("Vestas_V112_3MW", matrix=m, aggregate_time=None)
    assert b.dims == ("time", "dim_0")  # lazily loaded: the reference's dask branch order (aggregate.py:24-32)
    np.testing.assert_allclose(np.asarray(b.values), np.asarray(a.values).T, rtol=2e-5, atol=1e-6)
    assert lazy.data.largest_read <= 100
