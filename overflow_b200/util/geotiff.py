"""Minimal GeoTIFF / BigTIFF reader-writer with a GDAL-shaped surface (host I/O, pure numpy).

The reference does all raster I/O through GDAL (src/overflow/util/raster.py, flow_direction.py:103-120).
GDAL is not installed in this image, so `overflow_b200.util.raster` falls back to this module when
`osgeo` cannot be imported.  It covers what the D8 path needs and what GDAL's GTiff driver writes by
default: single-band, chunky, uncompressed, striped or tiled, little- or big-endian, classic or
BigTIFF.  Georeferencing (ModelPixelScale / ModelTiepoint / ModelTransformation / GeoKey tags) and
GDAL_NODATA are carried through verbatim.  Only the methods the path uses are provided:

    Dataset: RasterXSize, RasterYSize, GetRasterBand(1), GetProjection, SetProjection,
             GetGeoTransform, SetGeoTransform, FlushCache
    Band:    XSize, YSize, DataType, GetNoDataValue, SetNoDataValue, ReadAsArray, WriteArray
"""
import os
import struct

import numpy as np

# GDAL data type codes and names (gdal.GDT_*), as far as numpy can hold them
GDT = {"Byte": 1, "UInt16": 2, "Int16": 3, "UInt32": 4, "Int32": 5, "Float32": 6, "Float64": 7, "UInt64": 12,
       "Int64": 13, "Int8": 14}
GDT_NAME = {v: k for k, v in GDT.items()}
NP_OF_NAME = {"Byte": np.uint8, "UInt16": np.uint16, "Int16": np.int16, "UInt32": np.uint32, "Int32": np.int32,
              "Float32": np.float32, "Float64": np.float64, "UInt64": np.uint64, "Int64": np.int64, "Int8": np.int8}

_TAG_WIDTH, _TAG_HEIGHT, _TAG_BITS, _TAG_COMPRESSION, _TAG_PHOTOMETRIC = 256, 257, 258, 259, 262
_TAG_STRIP_OFFSETS, _TAG_SPP, _TAG_ROWS_PER_STRIP, _TAG_STRIP_BYTES, _TAG_PLANAR = 273, 277, 278, 279, 284
_TAG_TILE_W, _TAG_TILE_H, _TAG_TILE_OFFSETS, _TAG_TILE_BYTES, _TAG_SAMPLE_FORMAT = 322, 323, 324, 325, 339
_TAG_PIXEL_SCALE, _TAG_TIEPOINT, _TAG_TRANSFORM = 33550, 33922, 34264
_TAG_GEOKEYS, _TAG_GEO_DOUBLES, _TAG_GEO_ASCII, _TAG_GDAL_NODATA = 34735, 34736, 34737, 42113
_GEO_TAGS = (_TAG_PIXEL_SCALE, _TAG_TIEPOINT, _TAG_TRANSFORM, _TAG_GEOKEYS, _TAG_GEO_DOUBLES, _TAG_GEO_ASCII)

# TIFF field types: code -> (struct char, size)
_FT = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
       9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}


def _dtype_from_tags(bits, fmt, order):
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise ValueError(f"unsupported sample layout: {bits} bits per sample, sample format {fmt} "
                         "(the built-in reader handles 8/16/32/64-bit integers and 32/64-bit floats; use GDAL)")
    return np.dtype(f"{order}{kind}{bits // 8}")


def _name_of_dtype(dt):
    dt = np.dtype(dt)
    for name, npt in NP_OF_NAME.items():
        if np.dtype(npt) == dt.newbyteorder("="):
            return name
    raise TypeError(f"unsupported raster dtype {dt}")


class Band:
    def __init__(self, ds):
        self._ds = ds

    @property
    def XSize(self):
        return self._ds.RasterXSize

    @property
    def YSize(self):
        return self._ds.RasterYSize

    @property
    def DataType(self):
        return GDT[_name_of_dtype(self._ds._dtype)]

    def GetNoDataValue(self):
        return self._ds._nodata

    def SetNoDataValue(self, value):
        self._ds._set_nodata(value)

    def ReadAsArray(self, xoff=0, yoff=0, win_xsize=None, win_ysize=None, buf_obj=None):
        """The window as an array; with `buf_obj` (GDAL's name for it) the pixels go straight into that array --
        pinned host memory in the streaming drivers -- and it is returned."""
        win_xsize = self.XSize - xoff if win_xsize is None else win_xsize
        win_ysize = self.YSize - yoff if win_ysize is None else win_ysize
        return self._ds._read(int(xoff), int(yoff), int(win_xsize), int(win_ysize), buf_obj)

    def WriteArray(self, array, xoff=0, yoff=0):
        self._ds._write(np.asarray(array), int(xoff), int(yoff))

    def FlushCache(self):
        self._ds.FlushCache()


class Dataset:
    """One single-band GeoTIFF, memory-mapped."""

    def __init__(self):
        self.RasterXSize = self.RasterYSize = 0
        self.RasterCount = 1
        self._dtype = np.dtype(np.uint8)
        self._nodata = None
        self._geo = {}  # raw geo tags: tag -> (field type, values)
        self._mm = None  # numpy memmap of the pixel data when the layout is one contiguous run of rows
        self._chunks = None  # otherwise: (tile_w, tile_h, offsets) for tiled / non-contiguous files
        self._path = None
        self._writable = False
        self._nodata_slot = None

    # ---- GDAL-shaped surface
    def GetRasterBand(self, i):
        if i != 1:
            raise ValueError("only band 1 is supported")
        return Band(self)

    def GetProjection(self):
        """Opaque projection token: the raw GeoKey tags (round-trips through SetProjection)."""
        return {t: self._geo[t] for t in (_TAG_GEOKEYS, _TAG_GEO_DOUBLES, _TAG_GEO_ASCII) if t in self._geo}

    def SetProjection(self, proj):
        if isinstance(proj, dict):
            self._pending_geo.update(proj)
        elif proj:
            raise ValueError("the built-in GeoTIFF writer takes the projection token of GetProjection(), not WKT")

    def GetGeoTransform(self):
        if _TAG_TRANSFORM in self._geo:
            m = self._geo[_TAG_TRANSFORM][1]
            return (m[3], m[0], m[1], m[7], m[4], m[5])
        if _TAG_PIXEL_SCALE in self._geo and _TAG_TIEPOINT in self._geo:
            sx, sy = self._geo[_TAG_PIXEL_SCALE][1][:2]
            i, j, _, x, y, _ = self._geo[_TAG_TIEPOINT][1][:6]
            return (x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy)
        return (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)

    def SetGeoTransform(self, gt):
        x0, sx, rx, y0, ry, sy = [float(v) for v in gt]
        if rx == 0.0 and ry == 0.0 and sy <= 0.0 and tuple(gt) != (0.0, 1.0, 0.0, 0.0, 0.0, 1.0):
            self._pending_geo[_TAG_PIXEL_SCALE] = (12, (sx, -sy, 0.0))
            self._pending_geo[_TAG_TIEPOINT] = (12, (0.0, 0.0, 0.0, x0, y0, 0.0))
        elif tuple(gt) != (0.0, 1.0, 0.0, 0.0, 0.0, 1.0):
            self._pending_geo[_TAG_TRANSFORM] = (12, (sx, rx, 0.0, x0, ry, sy, 0.0, y0, 0, 0, 0, 0, 0, 0, 0, 1.0))

    def FlushCache(self):
        if self._writable and self._mm is not None:
            self._finalise_header()
            self._mm.flush()

    # ---- pixel access
    def _read(self, xoff, yoff, xs, ys, out=None):
        if xoff < 0 or yoff < 0 or xoff + xs > self.RasterXSize or yoff + ys > self.RasterYSize:
            raise ValueError("read window outside the raster")
        if out is not None:
            if out.shape != (ys, xs) or out.dtype != self._dtype.newbyteorder("="):
                raise ValueError("buf_obj must have the window's shape and the band's dtype")
            if self._mm is not None:
                out[...] = self._mm[yoff : yoff + ys, xoff : xoff + xs]
            else:
                out[...] = self._read(xoff, yoff, xs, ys)
            return out
        if self._mm is not None:
            return np.array(self._mm[yoff : yoff + ys, xoff : xoff + xs]).astype(self._dtype.newbyteorder("="), copy=False)
        tw, th, offsets = self._chunks[:3]
        strips = len(self._chunks) > 3  # strips: the last one holds only the rows that are left; tiles are always full
        out = np.empty((ys, xs), dtype=self._dtype.newbyteorder("="))
        across = (self.RasterXSize + tw - 1) // tw
        with open(self._path, "rb") as f:
            for ty in range(yoff // th, (yoff + ys - 1) // th + 1):
                rows_here = min(th, self.RasterYSize - ty * th) if strips else th
                for tx in range(xoff // tw, (xoff + xs - 1) // tw + 1):
                    f.seek(offsets[ty * across + tx])
                    raw = f.read(tw * rows_here * self._dtype.itemsize)
                    if len(raw) != tw * rows_here * self._dtype.itemsize:
                        raise ValueError(f"{self._path}: chunk ({ty}, {tx}) is truncated")
                    tile = np.frombuffer(raw, dtype=self._dtype).reshape(rows_here, tw)
                    y0, x0 = max(yoff, ty * th), max(xoff, tx * tw)
                    y1, x1 = min(yoff + ys, (ty + 1) * th), min(xoff + xs, (tx + 1) * tw)
                    out[y0 - yoff : y1 - yoff, x0 - xoff : x1 - xoff] = tile[y0 - ty * th : y1 - ty * th,
                                                                           x0 - tx * tw : x1 - tx * tw]
        return out

    def _write(self, array, xoff, yoff):
        if not self._writable:
            raise ValueError("dataset opened read-only")
        ys, xs = array.shape
        if xoff < 0 or yoff < 0 or xoff + xs > self.RasterXSize or yoff + ys > self.RasterYSize:
            raise ValueError("write window outside the raster")
        self._mm[yoff : yoff + ys, xoff : xoff + xs] = array

    # ---- nodata / header of files we create
    def _set_nodata(self, value):
        self._nodata = None if value is None else float(value)
        if self._writable:
            self._finalise_header()

    def _finalise_header(self):
        """(Re)write the tag block of a file created by create(): geo tags and GDAL_NODATA live after the pixels."""
        _write_header(self)


def _read_ifd(f, order, big):
    """Parse the first IFD: {tag: (field type, tuple of values)}."""
    if big:
        (n,) = struct.unpack(order + "Q", f.read(8))
        ent_size, cnt_fmt, inline = 20, "Q", 8
    else:
        (n,) = struct.unpack(order + "H", f.read(2))
        ent_size, cnt_fmt, inline = 12, "I", 4
    raw = f.read(n * ent_size)
    tags = {}
    for k in range(n):
        e = raw[k * ent_size : (k + 1) * ent_size]
        tag, ft = struct.unpack(order + "HH", e[:4])
        (count,) = struct.unpack(order + cnt_fmt, e[4 : 4 + inline])
        valraw = e[4 + inline :]
        if ft not in _FT:
            continue
        ch, size = _FT[ft]
        nbytes = size * count
        if nbytes > inline:
            (off,) = struct.unpack(order + cnt_fmt, valraw)
            pos = f.tell()
            f.seek(off)
            data = f.read(nbytes)
            f.seek(pos)
        else:
            data = valraw[:nbytes]
        if ft == 2:
            vals = (data.split(b"\0")[0].decode("latin-1"),)
        elif ft in (5, 10):
            nums = struct.unpack(order + ch[0] * (2 * count), data)
            vals = tuple(nums[i] / nums[i + 1] if nums[i + 1] else 0.0 for i in range(0, 2 * count, 2))
        else:
            vals = struct.unpack(order + ch * count, data)
        tags[tag] = (ft, vals)
    return tags


def open_geotiff(path, update=False):
    ds = Dataset()
    ds._path = path
    with open(path, "rb") as f:
        head = f.read(4)
        order = {b"II": "<", b"MM": ">"}.get(head[:2])
        if order is None:
            raise ValueError(f"{path}: not a TIFF file")
        (magic,) = struct.unpack(order + "H", head[2:4])
        if magic == 42:
            big = False
            (ifd,) = struct.unpack(order + "I", f.read(4))
        elif magic == 43:
            big = True
            f.read(4)
            (ifd,) = struct.unpack(order + "Q", f.read(8))
        else:
            raise ValueError(f"{path}: bad TIFF magic {magic}")
        f.seek(ifd)
        tags = _read_ifd(f, order, big)

    def one(tag, default=None):
        return tags[tag][1][0] if tag in tags else default

    if one(_TAG_COMPRESSION, 1) != 1:
        raise ValueError(f"{path}: compressed GeoTIFFs need GDAL (the built-in reader handles uncompressed files)")
    if one(_TAG_SPP, 1) != 1:
        raise ValueError(f"{path}: only single-band rasters are supported")
    if len(tags.get(_TAG_BITS, (3, (8,)))[1]) != 1:
        raise ValueError(f"{path}: only one sample per pixel is supported")
    ds.RasterXSize, ds.RasterYSize = int(one(_TAG_WIDTH)), int(one(_TAG_HEIGHT))
    ds._dtype = _dtype_from_tags(int(one(_TAG_BITS, 8)), int(one(_TAG_SAMPLE_FORMAT, 1)), order)
    if _TAG_GDAL_NODATA in tags:
        try:
            ds._nodata = float(tags[_TAG_GDAL_NODATA][1][0].strip())
        except ValueError:
            ds._nodata = None
    ds._geo = {t: tags[t] for t in _GEO_TAGS if t in tags}
    item = ds._dtype.itemsize
    if _TAG_TILE_OFFSETS in tags:
        ds._chunks = (int(one(_TAG_TILE_W)), int(one(_TAG_TILE_H)), tags[_TAG_TILE_OFFSETS][1])
    else:
        offsets = tags[_TAG_STRIP_OFFSETS][1]
        rps = int(one(_TAG_ROWS_PER_STRIP, ds.RasterYSize))
        row_bytes = ds.RasterXSize * item
        contiguous = all(offsets[i + 1] == offsets[i] + rps * row_bytes for i in range(len(offsets) - 1))
        if contiguous:
            ds._mm = np.memmap(path, dtype=ds._dtype, mode="r+" if update else "r", offset=offsets[0],
                               shape=(ds.RasterYSize, ds.RasterXSize))
            ds._writable = update
        else:
            ds._chunks = (ds.RasterXSize, rps, offsets, "strips")
    return ds


def _write_header(ds):
    """Classic TIFF when everything fits 32-bit offsets, BigTIFF otherwise.  Layout written by create():
    [header][pixels, one strip][IFD + out-of-line values]."""
    order = "<"
    item = ds._dtype.itemsize
    npix = ds.RasterXSize * ds.RasterYSize * item
    big = ds._big
    data_off = ds._data_off
    fmt = {"u": 1, "i": 2, "f": 3}[ds._dtype.kind]
    entries = [
        (_TAG_WIDTH, 4, (ds.RasterXSize,)), (_TAG_HEIGHT, 4, (ds.RasterYSize,)), (_TAG_BITS, 3, (item * 8,)),
        (_TAG_COMPRESSION, 3, (1,)), (_TAG_PHOTOMETRIC, 3, (1,)),
        (_TAG_STRIP_OFFSETS, 16 if big else 4, (data_off,)), (_TAG_SPP, 3, (1,)),
        (_TAG_ROWS_PER_STRIP, 4, (ds.RasterYSize,)), (_TAG_STRIP_BYTES, 16 if big else 4, (npix,)),
        (_TAG_PLANAR, 3, (1,)), (_TAG_SAMPLE_FORMAT, 3, (fmt,)),
    ]
    for tag, (ft, vals) in sorted(ds._pending_geo.items()):
        entries.append((tag, ft, tuple(vals)))
    if ds._nodata is not None:
        nd = ds._nodata
        txt = repr(int(nd)) if float(nd).is_integer() and abs(nd) < 1e15 else repr(float(nd))
        entries.append((_TAG_GDAL_NODATA, 2, (txt,)))
    entries.sort(key=lambda e: e[0])
    ifd_off = data_off + npix + (-(data_off + npix)) % 16
    ent_size, inline, cnt_fmt = (20, 8, "Q") if big else (12, 4, "I")
    head_size = (8 if big else 2) + len(entries) * ent_size + (8 if big else 4)
    extra_off = ifd_off + head_size
    body, extra = b"", b""
    for tag, ft, vals in entries:
        if ft == 2:
            data = vals[0].encode("latin-1") + b"\0"
            count = len(data)
        else:
            ch, _ = _FT[ft]
            data = struct.pack(order + ch * len(vals), *vals)
            count = len(vals)
        if len(data) <= inline:
            field = data.ljust(inline, b"\0")
        else:
            field = struct.pack(order + cnt_fmt, extra_off + len(extra))
            extra += data + b"\0" * ((-len(data)) % 2)
        body += struct.pack(order + "HH" + cnt_fmt, tag, ft, count) + field
    ifd = (struct.pack(order + ("Q" if big else "H"), len(entries)) + body + struct.pack(order + ("Q" if big else "I"), 0))
    with open(ds._path, "r+b") as f:
        if big:
            f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off))
        else:
            f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        f.seek(ifd_off)
        f.write(ifd + extra)
        f.truncate()


def create_geotiff(path, xsize, ysize, dtype_name):
    """New single-band uncompressed GeoTIFF, pixels memory-mapped for windowed writes."""
    ds = Dataset()
    ds._path = path
    ds.RasterXSize, ds.RasterYSize = int(xsize), int(ysize)
    ds._dtype = np.dtype(NP_OF_NAME[dtype_name]).newbyteorder("<")
    ds._pending_geo = {}
    npix = ds.RasterXSize * ds.RasterYSize * ds._dtype.itemsize
    ds._big = npix + 65536 >= (1 << 32)
    ds._data_off = 16
    with open(path, "wb") as f:
        f.truncate(ds._data_off + npix)
    ds._writable = True
    if npix:
        ds._mm = np.memmap(path, dtype=ds._dtype, mode="r+", offset=ds._data_off, shape=(ds.RasterYSize, ds.RasterXSize))
    else:
        ds._mm = np.zeros((ds.RasterYSize, ds.RasterXSize), dtype=ds._dtype)
    _write_header(ds)
    return ds


def exists(path):
    return os.path.exists(path)
