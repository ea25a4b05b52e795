"""Regenerates tests/golden/prs_clumping.npz from the reference's own RDS fixtures (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz).

  tests/testthat/testdata/pval.rds      4,542 doubles: predict(gwas, log10 = FALSE) of tests/testthat/test-6-PRS.R:19-22
  tests/testthat/testdata/clumping.rds  sorted 1-based indices snp_clumping() kept in tests/testthat/test-6-PRS.R:25-30

RDS = gzip stream of R's XDR serialisation: "X\n", three int32 (format version 2, writer, min reader), then one item:
int32 flags (low byte = SEXP type: 13 INTSXP, 14 REALSXP; no attributes in these two files), int32 length, big-endian
payload.  Only that subset is parsed -- anything else raises.
"""
import gzip
import os
import struct
import sys

import numpy as np

REF = os.environ.get("BIGSNPR_REFERENCE", "/root/reference")


def read_rds_vector(path):
    d = gzip.decompress(open(path, "rb").read())
    if d[:2] != b"X\n":
        raise ValueError("not an XDR serialisation: %r" % d[:2])
    version, _writer, _minreader = struct.unpack(">3i", d[2:14])
    if version != 2:
        raise ValueError("serialisation version %d not handled" % version)
    flags, length = struct.unpack(">2i", d[14:22])
    sxp, has_attr = flags & 0xFF, bool(flags & 0x200)
    if has_attr:
        raise ValueError("attributes not handled")
    if sxp == 13:
        out = np.frombuffer(d, dtype=">i4", count=length, offset=22).astype(np.int32)
        end = 22 + 4 * length
    elif sxp == 14:
        out = np.frombuffer(d, dtype=">f8", count=length, offset=22).astype(np.float64)
        end = 22 + 8 * length
    else:
        raise ValueError("SEXP type %d not handled" % sxp)
    if end != len(d):
        raise ValueError("trailing bytes")
    return out


def main():
    td = os.path.join(REF, "tests", "testthat", "testdata")
    pval = read_rds_vector(os.path.join(td, "pval.rds"))
    keep = read_rds_vector(os.path.join(td, "clumping.rds"))
    assert pval.dtype == np.float64 and keep.dtype == np.int32
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prs_clumping.npz")
    np.savez_compressed(out, pval=pval, keep=keep)
    print("wrote", out, "pval", pval.shape, float(pval.min()), float(pval.max()), "keep", keep.shape, int(keep.min()),
          int(keep.max()))


if __name__ == "__main__":
    sys.exit(main())
