"""Writes Adam7-interlaced (and plain) PNG files of every colour type / bit depth with the five filter types, for
tests/test_image_decode.py (PIL and OpenCV cannot write interlaced PNGs).  RFC 2083 only: zlib + struct + numpy."""
import struct
import zlib

import numpy as np

ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def _chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def _pack_rows(samples, depth):
    """samples: (h, w, channels) integer array of `depth`-bit values -> list of packed scanlines (bytes)"""
    h, w, ch = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1).astype(np.uint32)
        if depth == 8:
            rows.append(flat.astype(np.uint8).tobytes())
        elif depth == 16:
            rows.append(flat.astype(">u2").tobytes())
        else:
            per = 8 // depth
            pad = (-len(flat)) % per
            f = np.concatenate([flat, np.zeros(pad, np.uint32)]).reshape(-1, per)
            shifts = np.arange(per - 1, -1, -1) * depth
            rows.append((f << shifts).sum(axis=1).astype(np.uint8).tobytes())
    return rows


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def _filter(rows, bpp, first_filter):
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, row in enumerate(rows):
        ft = (first_filter + y) % 5
        out.append(ft)
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = [0, a, b, (a + b) >> 1, _paeth(a, b, c)][ft]
            out.append((v - pred) & 255)
        prev = row
    return bytes(out)


def write_png(path, samples, colour, depth, interlace, palette=None, trns=None):
    """samples: (h, w, channels) array of raw sample values (palette indices for colour type 3)"""
    h, w, ch = samples.shape
    assert ch == CHANNELS[colour]
    bpp = max(1, ch * depth // 8)
    data = bytearray()
    passes = ADAM7 if interlace else [(0, 0, 1, 1)]
    for n, (x0, y0, dx, dy) in enumerate(passes):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        data += _filter(_pack_rows(sub, depth), bpp, n)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += _chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        out += _chunk(b"tRNS", bytes(trns))
    z = zlib.compress(bytes(data), 6)
    half = len(z) // 2
    out += _chunk(b"IDAT", z[:half]) + _chunk(b"IDAT", z[half:]) + _chunk(b"IEND", b"")
    open(path, "wb").write(out)


def cases():
    """(name, colour, depth, has palette, trns)"""
    return [("g1", 0, 1), ("g2", 0, 2), ("g4", 0, 4), ("g8", 0, 8), ("g16", 0, 16), ("rgb8", 2, 8), ("rgb16", 2, 16),
            ("p1", 3, 1), ("p2", 3, 2), ("p4", 3, 4), ("p8", 3, 8), ("ga8", 4, 8), ("ga16", 4, 16), ("rgba8", 6, 8), ("rgba16", 6, 16)]


def make(path, name, colour, depth, w, h, interlace, seed):
    rng = np.random.default_rng(seed)
    ch = CHANNELS[colour]
    samples = rng.integers(0, 1 << depth, (h, w, ch))
    palette = trns = None
    if colour == 3:
        palette = rng.integers(0, 256, (1 << depth, 3))
        trns = rng.integers(0, 256, max(1, (1 << depth) // 2)).astype(np.uint8).tolist()
    elif colour == 0 and depth <= 8:
        trns = [0, int(samples[0, 0, 0])]
    elif colour == 2 and depth == 8:
        trns = [0, int(samples[0, 0, 0]), 0, int(samples[0, 0, 1]), 0, int(samples[0, 0, 2])]
    write_png(path, samples, colour, depth, interlace, palette, trns)
