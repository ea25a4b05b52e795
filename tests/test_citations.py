"""Every `file:line` citation into the reference tree (C-ABI header, oracle, CUDA sources, host mirror, docs) must
point at an existing file and lines inside it.  Runs only where the reference checkout is mounted (this container);
skipped on the GPU box, where /root/reference does not exist."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"\b((?:src|R|tests/testthat|inst/extdata)/[A-Za-z0-9_./-]+\.(?:cpp|h|R|ld|rds|bed))(?::(\d+)(?:-(\d+))?)?")

FILES = ["include/bsgpu.h", "oracle/bsg_oracle.c", "oracle/ref.py", "bigsnpr_b200/api.py", "bigsnpr_b200/dist.py",
         "r_shim/bigsnpr_shim.c", "INTEGRATION.md", "DESIGN.md"] + [
    os.path.join("bigsnpr_b200/csrc", f) for f in sorted(os.listdir(os.path.join(ROOT, "bigsnpr_b200", "csrc")))
    if f.endswith((".cu", ".cuh"))] + [os.path.join("tests", f) for f in sorted(os.listdir(os.path.join(ROOT, "tests")))
                                        if f.endswith(".py") and f != "test_citations.py"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_reference_citations_resolve():
    nlines, bad, total = {}, [], 0
    for rel in FILES:
        text = open(os.path.join(ROOT, rel), errors="replace").read()
        for m in CITE.finditer(text):
            path, a, b = m.group(1), m.group(2), m.group(3)
            full = os.path.join(REF, path)
            total += 1
            if not os.path.isfile(full):
                bad.append((rel, m.group(0), "no such file"))
                continue
            if a is None:
                continue
            if full not in nlines:
                with open(full, "rb") as f:
                    nlines[full] = sum(1 for _ in f)
            lo, hi = int(a), int(b or a)
            if not (1 <= lo <= hi <= nlines[full]):
                bad.append((rel, m.group(0), "file has %d lines" % nlines[full]))
    assert total > 200, total
    assert not bad, bad[:20]
