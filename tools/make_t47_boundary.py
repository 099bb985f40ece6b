"""Synthesise T47 (144x72) boundary data from the packed T30 reference files (BASELINE configs[3];
the reference ships T30 data only).

Every field is resampled by NEAREST NEIGHBOUR in (latitude, longitude) from the T30 grid: that
keeps the land-sea mask, the fill values (9.96921e36 over masked points, input_output.f90:39) and
the fields mutually consistent, so the loaders' flip / fill / forchk logic (boundaries.f90:28-68)
sees the same kind of data as at T30.  Latitudes are the reference's approximate Gaussian
latitudes (geometry.f90:68, evaluated in real32), file order N->S as in the reference files.
Output format: the SPDYBC01 pack of tools/pack_boundary.py with ix=144, il=72; `ssta` is cut to
NSSTA months to keep the file small.  numpy only — runs on the GPU box too.

usage: python tools/make_t47_boundary.py [data/bc_t30.bin] [data/bc_t47.bin]
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSSTA = 48   # 1979-01 .. 1982-12: covers a run that starts 1982-01-01 (record = (year-1979)*12 + month, sea_model.f90:377)


def gauss_lat_deg(il):
    """degrees, S->N, from sia_half(j) = cos(3.141592654*(j - 0.25)/(il + 0.5)) in real32 (geometry.f90:68)"""
    iy = il // 2
    j = np.arange(1, iy + 1, dtype=np.float32)
    sia_half = np.cos(np.float32(3.141592654) * (j - np.float32(0.25)) / (np.float32(il) + np.float32(0.5))).astype(np.float32)
    south = -np.degrees(np.arcsin(sia_half.astype(np.float64)))       # j = 1 southernmost
    return np.concatenate([south, -south[::-1]])


def read_pack(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"SPDYBC01", "not a SPDYBC01 pack"
    ix, il, nf = struct.unpack_from("<iii", raw, 8)
    off, fields = 20, []
    for _ in range(nf):
        name = raw[off:off + 16].rstrip(b"\0").decode()
        nrec, = struct.unpack_from("<i", raw, off + 16)
        a = np.frombuffer(raw, "<f4", nrec * il * ix, off + 20).reshape(nrec, il, ix)
        fields.append((name, a))
        off += 20 + 4 * nrec * il * ix
    return ix, il, fields


def write_pack(path, ix, il, fields):
    with open(path, "wb") as f:
        f.write(b"SPDYBC01")
        f.write(struct.pack("<iii", ix, il, len(fields)))
        for name, a in fields:
            f.write(name.encode().ljust(16, b"\0"))
            f.write(struct.pack("<i", a.shape[0]))
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())


def make(src, dst, ix2=144, il2=72, nssta=NSSTA):
    ix1, il1, fields = read_pack(src)
    lat1 = gauss_lat_deg(il1)[::-1]       # file order N->S
    lat2 = gauss_lat_deg(il2)[::-1]
    jj = np.abs(lat2[:, None] - lat1[None, :]).argmin(axis=1)
    lon1 = np.arange(ix1) * (360.0 / ix1)
    lon2 = np.arange(ix2) * (360.0 / ix2)
    d = np.abs(lon2[:, None] - lon1[None, :])
    ii = np.minimum(d, 360.0 - d).argmin(axis=1)
    out = []
    for name, a in fields:
        if name == "ssta":
            a = a[:nssta]
        out.append((name, a[:, jj][:, :, ii]))
    write_pack(dst, ix2, il2, out)
    return dst


def ensure(dst=None):
    """create data/bc_t47.bin if it is missing (tests and bench call this; the file is git-ignored)"""
    dst = dst or os.path.join(ROOT, "data", "bc_t47.bin")
    if not os.path.exists(dst):
        tmp = dst + f".tmp{os.getpid()}"
        make(os.path.join(ROOT, "data", "bc_t30.bin"), tmp)
        os.replace(tmp, dst)
    return dst


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "data", "bc_t30.bin")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "data", "bc_t47.bin")
    print("wrote", make(src, dst), os.path.getsize(dst), "bytes")
