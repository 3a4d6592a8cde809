"""Every ``file.rs:line`` citation of the reference in the header, the docs, the oracle, the transcripts and the product
sources points at an existing file of /root/reference and at lines that file has.  Skipped where the reference is not
mounted (the GPU box)."""
import collections
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
REF = pathlib.Path("/root/reference")

PAT = re.compile(r"([A-Za-z0-9_\-./…]*?([A-Za-z0-9_\-]+\.(?:rs|toml|cu)))`?:(\d+)(?:-(\d+))?((?:,\s*\d+(?:-\d+)?)*)")


@pytest.mark.skipif(not REF.exists(), reason="the reference is only mounted in the build container")
def test_reference_citations_resolve():
    files = collections.defaultdict(list)
    for p in REF.rglob("*"):
        if p.is_file() and p.suffix in (".rs", ".toml", ".cu"):
            files[p.name].append(p)
    lens = {}

    def n_lines(p):
        if p not in lens:
            lens[p] = sum(1 for _ in open(p, errors="ignore"))
        return lens[p]

    own = {p.name for p in ROOT.rglob("*") if p.is_file() and ".git" not in p.parts}
    srcs = [ROOT / "include" / "crb.h", ROOT / "DESIGN.md", ROOT / "INTEGRATION.md", ROOT / "README.md"]
    for pattern in ("oracle/*.cpp", "oracle/*.hpp", "oracle/*.py", "tests/*.py", "tests/*.c", "cellular_raza_b200/*.py",
                    "cellular_raza_b200/csrc/*", "rust/**/*.rs", "bench.py"):
        srcs += list(ROOT.glob(pattern))
    bad, total = [], 0
    for s in srcs:
        for m in PAT.finditer(s.read_text(errors="ignore")):
            full, name = m.group(1), m.group(2)
            nums = [int(m.group(3))] + ([int(m.group(4))] if m.group(4) else []) + [int(x) for x in re.findall(r"\d+", m.group(5))]
            cands = files.get(name)
            if not cands:
                if name not in own:
                    bad.append((s.name, full, "no such file in the reference"))
                continue
            frag = [f for f in re.split(r"[/…]+", full) if f and f != name]
            cs = [c for c in cands if all(f in str(c) for f in frag)] or cands
            total += 1
            if all(n_lines(c) < max(nums) for c in cs):
                bad.append((s.name, full, f"line {max(nums)} beyond the end of the file"))
    assert total > 300, total
    assert not bad, bad[:10]
