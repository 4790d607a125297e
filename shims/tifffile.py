"""Baseline TIFF reader / writer for the reference's dataset files ({id}_lr.tif [h,w,B], {id}_pan.tif [H,W],
{id}_mul.tif [H,W,B], uint16) — stand-in for tifffile.imread (dataset/utils.py:37) in an image without tifffile.
Supports what the fixtures and gdal-shim outputs use: little- or big-endian classic TIFF, uncompressed, strips,
chunky (PlanarConfiguration 1) or planar (2, returned as [C,H,W]) samples of uint8 / uint16 / uint32 / float32."""
import struct

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 16: ("Q", 8)}


def imread(path):
    with open(path, "rb") as f:
        buf = f.read()
    bo = {b"II": "<", b"MM": ">"}.get(buf[:2])
    if bo is None or struct.unpack(bo + "H", buf[2:4])[0] != 42:
        raise ValueError(f"{path}: not a classic TIFF file")
    off = struct.unpack(bo + "I", buf[4:8])[0]
    n = struct.unpack(bo + "H", buf[off:off + 2])[0]
    tags = {}
    for i in range(n):
        e = off + 2 + 12 * i
        tag, typ, cnt = struct.unpack(bo + "HHI", buf[e:e + 8])
        if typ not in _TYPES or typ == 5:
            continue
        code, size = _TYPES[typ]
        raw = buf[e + 8:e + 12] if size * cnt <= 4 else None
        if raw is None:
            p = struct.unpack(bo + "I", buf[e + 8:e + 12])[0]
            raw = buf[p:p + size * cnt]
        tags[tag] = struct.unpack(bo + code * cnt, raw[:size * cnt]) if typ != 2 else raw
    w, h = tags[256][0], tags[257][0]
    spp = tags.get(277, (1,))[0]
    bits = tags.get(258, (8,))[0]
    fmt = tags.get(339, (1,))[0]
    if tags.get(259, (1,))[0] != 1:
        raise ValueError(f"{path}: compressed TIFF is not supported by this shim")
    planar = tags.get(284, (1,))[0]
    dt = {(8, 1): "u1", (16, 1): "u2", (32, 1): "u4", (32, 3): "f4", (16, 2): "i2", (64, 3): "f8"}[(bits, fmt)]
    data = b"".join(buf[o:o + c] for o, c in zip(tags[273], tags[279]))
    arr = np.frombuffer(data, dtype=np.dtype(dt).newbyteorder(bo))
    if spp == 1:
        return arr[:h * w].reshape(h, w).astype(dt)
    if planar == 2:
        return arr[:spp * h * w].reshape(spp, h, w).astype(dt)
    return arr[:h * w * spp].reshape(h, w, spp).astype(dt)


def imwrite(path, array, planar=False):
    """[H,W] or [H,W,C] (chunky; planar=True: [C,H,W] written as separate planes)."""
    a = np.ascontiguousarray(array)
    if a.dtype not in (np.uint8, np.uint16, np.uint32, np.float32):
        raise TypeError(f"unsupported dtype {a.dtype}")
    if a.ndim == 2:
        h, w, spp = a.shape[0], a.shape[1], 1
    elif planar:
        spp, h, w = a.shape
    else:
        h, w, spp = a.shape
    bits = a.dtype.itemsize * 8
    fmt = 3 if a.dtype == np.float32 else 1
    data = a.astype(a.dtype.newbyteorder("<")).tobytes()
    nstrips = spp if (planar and a.ndim == 3) else 1
    strip = len(data) // nstrips
    entries = []            # (tag, type, count, values)
    entries.append((256, 4, 1, [w]))
    entries.append((257, 4, 1, [h]))
    entries.append((258, 3, spp, [bits] * spp))
    entries.append((259, 3, 1, [1]))
    entries.append((262, 3, 1, [1]))
    entries.append((273, 4, nstrips, None))          # strip offsets, patched below
    entries.append((277, 3, 1, [spp]))
    entries.append((278, 4, 1, [h]))
    entries.append((279, 4, nstrips, [strip] * nstrips))
    entries.append((284, 3, 1, [2 if nstrips > 1 else 1]))
    if spp > 1:
        entries.append((338, 3, spp - 1, [0] * (spp - 1)))
    entries.append((339, 3, spp, [fmt] * spp))
    entries.sort(key=lambda e: e[0])
    ifd_off = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_size
    extra = b""
    size_of = {3: 2, 4: 4}
    code_of = {3: "H", 4: "I"}
    # first pass: where does the pixel data start
    ext_total = sum(size_of[t] * c for _, t, c, _ in entries if size_of[t] * c > 4)
    data_off = extra_off + ext_total
    data_off += data_off & 1
    body = struct.pack("<H", len(entries))
    for tag, typ, cnt, vals in entries:
        if tag == 273:
            vals = [data_off + i * strip for i in range(nstrips)]
        raw = struct.pack("<" + code_of[typ] * cnt, *vals)
        if len(raw) <= 4:
            body += struct.pack("<HHI", tag, typ, cnt) + raw.ljust(4, b"\0")
        else:
            body += struct.pack("<HHII", tag, typ, cnt, extra_off + len(extra))
            extra += raw
    body += struct.pack("<I", 0)
    head = b"II" + struct.pack("<HI", 42, ifd_off)
    blob = head + body + extra
    blob = blob.ljust(data_off, b"\0")
    with open(path, "wb") as f:
        f.write(blob + data)


imsave = imwrite
