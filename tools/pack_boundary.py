"""Pack the reference's T30 boundary files (data/bc/t30/{clim,anom}/*.nc, NetCDF-4/HDF5 with
contiguous little-endian float32 variables) into one flat file that travels with the repo
(the GPU box has no /root/reference and no HDF5 library).

Layout of data/bc_t30.bin (little endian):
    8 bytes  magic "SPDYBC01"
    int32    ix, il, nfields
    per field: char name[16] (NUL padded), int32 nrec, float32 data[nrec][il][ix]
The data are the files' own values in the files' own order (latitude N->S, no fill-value
handling): flipping and missing-value logic stay in the loaders (input_output.f90:23-92).
Dataset byte offsets were found by walking the HDF5 object headers and are
guarded by file size + sha256 prefix.  `ssta` is cut to the first NSSTA months (1979-01 ...).

usage: python tools/pack_boundary.py [/root/reference/data/bc/t30] [out.bin] [--months N]
       --months N   SST-anomaly records to keep (default 72 = 1979-01 .. 1984-12, the shipped data/bc_t30.bin; the file holds 420)
"""
import hashlib
import os
import struct
import sys

import numpy as np

IX, IL = 96, 48
NSSTA = 72   # 1979-01 .. 1984-12
FILES = {
    "clim/surface.nc": (107750, "858ecc94f283adbc", [("orog", 6144, 1), ("lsm", 28902, 1), ("alb", 67034, 1), ("vegh", 48058, 1), ("vegl", 89318, 1)]),
    "clim/land.nc": (231539, "6139d2f8db263ee8", [("stl", 6454, 12)]),
    "clim/sea_surface_temperature.nc": (231539, "247b7f2e487ec73b", [("sst", 6454, 12)]),
    "clim/sea_ice.nc": (231541, "b37daa33189c0788", [("icec", 6454, 12)]),
    "clim/snow.nc": (231414, "e49a07f10936e936", [("snowd", 6454, 12)]),
    "clim/soil.nc": (675398, "49ca070b0e7c43ce", [("swl1", 6454, 12), ("swl2", 232240, 12)]),
    "anom/sea_surface_temperature_anomaly.nc": (7755157, "e6572d345de3b8de", [("ssta", 6454, NSSTA)]),
}


def main():
    global NSSTA
    if "--months" in sys.argv:
        i = sys.argv.index("--months")
        NSSTA = int(sys.argv[i + 1])
        del sys.argv[i:i + 2]
        assert 1 <= NSSTA <= 420
        FILES["anom/sea_surface_temperature_anomaly.nc"][2][0] = ("ssta", 6454, NSSTA)
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/bc/t30"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "bc_t30.bin")
    fields = []
    for rel, (size, sha, vars_) in FILES.items():
        path = os.path.join(src, rel)
        raw = open(path, "rb").read()
        assert len(raw) == size, (rel, len(raw))
        assert hashlib.sha256(raw).hexdigest().startswith(sha), rel
        for name, off, nrec in vars_:
            a = np.frombuffer(raw, dtype="<f4", count=nrec * IL * IX, offset=off).reshape(nrec, IL, IX)
            fields.append((name, a))
    with open(out, "wb") as f:
        f.write(b"SPDYBC01")
        f.write(struct.pack("<iii", IX, IL, len(fields)))
        for name, a in fields:
            f.write(name.encode().ljust(16, b"\0"))
            f.write(struct.pack("<i", a.shape[0]))
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
    print("wrote", out, os.path.getsize(out), "bytes;", ", ".join(f"{n}{list(a.shape)}" for n, a in fields))
    for n, a in fields:
        ok = a[np.abs(a) < 1e30]
        print(f"  {n:6s} min {ok.min():12.4f} max {ok.max():12.4f} nfill {(np.abs(a) >= 1e30).sum()}")


if __name__ == "__main__":
    main()
