"""Minimal OpenEXR scan-line reader / writer for the baked shading maps (numpy + zlib only).

The reference writes its maps with `cv2.imwrite(path + '.exr', img[:, :, [2, 1, 0]])` (bake_shading.py:131,202-203) and reads them
back with `cv2.imread(path, -1)[..., [2, 1, 0]]` (utils/dataset/*: the `diffuse/` and `specular/` folders).  OpenCV only does that
when it was built with OpenEXR and `OPENCV_IO_ENABLE_OPENEXR=1` is set; this module removes the dependency: `write_exr` /
`read_exr` move (H, W, C) float32 arrays to and from single-part scan-line files -- uncompressed or ZIP (16 scan lines per
block, OpenCV's default), FLOAT or HALF channels -- which is what either side produces.  Channel order in the file is
alphabetical (the format's rule); arrays are RGB in, RGB out.  tests/test_exr_cpu.py cross-checks both directions with OpenCV.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_MAGIC = 20000630
_PT = {0: np.dtype("<u4"), 1: np.dtype("<f2"), 2: np.dtype("<f4")}
_LINES = {0: 1, 2: 1, 3: 16}     # NO_COMPRESSION, ZIPS, ZIP


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def _zip_pack(raw):
    """OpenEXR's ZIP pre-filter: de-interleave even/odd bytes, then byte-wise delta, then deflate."""
    a = np.frombuffer(raw, np.uint8)
    t = np.concatenate([a[0::2], a[1::2]]).astype(np.int16)
    d = t.copy()
    d[1:] = t[1:] - t[:-1] + 128
    out = zlib.compress((d & 0xFF).astype(np.uint8).tobytes(), 6)
    return out if len(out) < len(raw) else raw


def _zip_unpack(data, size):
    if len(data) == size:                                      # blocks that do not shrink are stored raw
        return data
    d = np.frombuffer(zlib.decompress(data), np.uint8).astype(np.int64)
    d[1:] -= 128
    t = (np.cumsum(d) & 0xFF).astype(np.uint8)
    half = (size + 1) // 2
    out = np.empty(size, np.uint8)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def write_exr(path, img, channels="RGB", half=False, compression="zip"):
    """img: (H, W, C) or (H, W) array -> `path`.  channels names the C planes (default R, G, B); half=True stores fp16."""
    img = np.asarray(img, np.float32)
    if img.ndim == 2:
        img, channels = img[:, :, None], "Y"
    H, W, C = img.shape
    names = list(channels)[:C] if C <= len(channels) else ["C%d" % i for i in range(C)]
    order = sorted(range(C), key=lambda i: names[i])           # channels are stored alphabetically
    pt = 1 if half else 2
    comp = {"none": 0, "zips": 2, "zip": 3}[compression]
    ch = b"".join(names[i].encode() + b"\0" + struct.pack("<iBBBBii", pt, 0, 0, 0, 0, 1, 1) for i in order) + b"\0"
    box = struct.pack("<iiii", 0, 0, W - 1, H - 1)
    header = b"".join([
        _attr("channels", "chlist", ch), _attr("compression", "compression", bytes([comp])), _attr("dataWindow", "box2i", box),
        _attr("displayWindow", "box2i", box), _attr("lineOrder", "lineOrder", b"\0"), _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)), _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))]) + b"\0"
    lines = _LINES[comp]
    blocks = []
    for y0 in range(0, H, lines):
        rows = img[y0:y0 + lines][:, :, order].astype(_PT[pt])   # (h, W, C) -> per scan line: channel planes one after the other
        raw = np.ascontiguousarray(rows.transpose(0, 2, 1)).tobytes()
        blocks.append((y0, _zip_pack(raw) if comp else raw))
    head = struct.pack("<ii", _MAGIC, 2) + header
    off = len(head) + 8 * len(blocks)
    table, body = [], []
    for y0, data in blocks:
        table.append(struct.pack("<Q", off))
        body.append(struct.pack("<ii", y0, len(data)) + data)
        off += 8 + len(data)
    with open(path, "wb") as f:
        f.write(head + b"".join(table) + b"".join(body))


def read_exr(path, channels=None):
    """-> (H, W, C) float32.  channels: names to return, in that order (default: R, G, B if present, else the file's order)."""
    buf = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", buf, 0)
    if magic != _MAGIC:
        raise ValueError("%s: not an OpenEXR file" % path)
    if version & 0x1a00:
        raise ValueError("%s: tiled / multi-part / deep files are not supported" % path)
    pos, attrs = 8, {}
    while buf[pos] != 0:
        e = buf.index(b"\0", pos)
        name = buf[pos:e].decode()
        e2 = buf.index(b"\0", e + 1)
        size = struct.unpack_from("<i", buf, e2 + 1)[0]
        attrs[name] = buf[e2 + 5:e2 + 5 + size]
        pos = e2 + 5 + size
    pos += 1
    chl, p, file_ch = attrs["channels"], 0, []
    while chl[p] != 0:
        e = chl.index(b"\0", p)
        pt, xs, ys = struct.unpack_from("<i4xii", chl, e + 1)
        if xs != 1 or ys != 1:
            raise ValueError("%s: sub-sampled channels are not supported" % path)
        file_ch.append((chl[p:e].decode(), pt))
        p = e + 17
    comp = attrs["compression"][0]
    if comp not in _LINES:
        raise ValueError("%s: compression %d is not supported (none / zips / zip are)" % (path, comp))
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"])
    W, H = x1 - x0 + 1, y1 - y0 + 1
    lines = _LINES[comp]
    n_blocks = (H + lines - 1) // lines
    offsets = struct.unpack_from("<%dQ" % n_blocks, buf, pos)
    planes = {n: np.empty((H, W), np.float32) for n, _ in file_ch}
    row_bytes = sum(_PT[pt].itemsize for _, pt in file_ch) * W
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        h = min(lines, y1 - y + 1)
        raw = buf[off + 8:off + 8 + size]
        if comp:
            raw = _zip_unpack(raw, h * row_bytes)
        q = 0
        for r in range(h):
            for n, pt in file_ch:
                dt = _PT[pt]
                planes[n][y - y0 + r] = np.frombuffer(raw, dt, W, q)
                q += dt.itemsize * W
    if channels is None:
        channels = "RGB" if all(c in planes for c in "RGB") else [n for n, _ in file_ch]
    return np.stack([planes[c] for c in channels], -1)
