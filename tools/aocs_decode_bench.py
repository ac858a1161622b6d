#!/usr/bin/env python3
"""Throughput of the on-device AOCS decoder (k_aocs_decode): tiles the reference-written golden column files
(tests/golden/aocs_columns.npz) to a few hundred MB each and reports the kernel's own time (launch trace), file bytes/s
and rows/s per column kind.  Profiling aid; the H2D copy of the file (pageable host memory here) is not in the number."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cloudberry_b200 import capi  # noqa: E402
from test_aocs_format import CASES, ZCASES, ZSTDCASES  # noqa: E402
from test_gpu_aocs import DECODE  # noqa: E402
from oracle import aocs_format as A  # noqa: E402


def main():
    ctx = capi.Context(0)
    target = 256 << 20
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for case in CASES + ZCASES + ZSTDCASES:
        if only and not case[0].startswith(tuple(only.split(","))):
            continue
        name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls = case[:9]
        ctype_z = 0 if len(case) <= 9 else 2 if name.startswith("zstd") else 1
        if len(values) < 1000:
            continue
        k = max(1, target // len(raw))
        content = sum(b["dlen"] for b in A.walk_blocks_ex(raw, checksum))
        k = max(1, min(k, (1 << 30) // max(len(values), 1), (3 << 30) // max(content, 1)))     # bound rows and inflated bytes
        big = raw * k
        n = len(values) * k
        ctype, attlen, varkind, align = DECODE[typname]
        rel = capi.DeviceRelation(ctx, n, [ctype], dscales=[dscale])
        rel.load_aocs_column(0, big, checksum, attlen, varkind, align, compresstype=ctype_z)      # warm-up
        ctx.trace_begin()
        got = rel.load_aocs_column(0, big, checksum, attlen, varkind, align, compresstype=ctype_z)
        tr = ctx.trace_end()
        ms = sum(m for nme, m in tr if nme == "k_aocs_decode")
        vms = sum(m for nme, m in tr if nme == "k_aocs_verify")
        ims = sum(m for nme, m in tr if nme in ("k_aocs_inflate", "k_aocs_unzstd"))
        cpu = ""
        if ctype_z:
            # the library the reference calls, one host core, on this column's own compressed blocks (uncompress / ZSTD_decompress)
            import time
            import zlib
            blocks = [(raw[b["off"]:b["off"] + b["clen"]], b["dlen"]) for b in A.walk_blocks_ex(raw, checksum) if b["clen"]]
            if blocks:
                if ctype_z == 2:
                    import pyarrow as pa
                    codec = pa.Codec("zstd")
                    dec = lambda z, n: codec.decompress(z, decompressed_size=n, asbytes=True)      # noqa: E731
                else:
                    dec = lambda z, n: zlib.decompress(z)                                            # noqa: E731
                t0, done = time.perf_counter(), 0
                while time.perf_counter() - t0 < 0.2:
                    for z, n_out in blocks:
                        dec(z, n_out)
                        done += n_out
                cpu = "  cpu core %.2f GB/s out" % (done / (time.perf_counter() - t0) / 1e9)
        assert got == n
        content = sum(b["dlen"] for b in A.walk_blocks_ex(raw, checksum) if b["clen"]) * k
        print("%-30s %8.1f MB file  %10d rows  kernel %7.3f ms  %7.1f GB/s of file  %7.2f G rows/s  (%d blocks)  crc32c %s  inflate %s" %
              (name, len(big) / 1e6, n, ms, len(big) / ms / 1e6, n / ms / 1e6, nblocks * k,
               "%.3f ms %.1f GB/s" % (vms, len(big) / vms / 1e6) if vms else "off",
               ("%.3f ms %.1f GB/s out" % (ims, content / ims / 1e6) if ims else "-") + cpu))
        rel.free()
    ctx.close()


if __name__ == "__main__":
    main()
