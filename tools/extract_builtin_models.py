#!/usr/bin/env python3
"""Regenerate phylocsfpp_b200/data/ from the reference's embedded model tables.

The reference ships its 11 built-in parameter sets (species tree, coding and
non-coding empirical codon model) as C arrays in src/models.hpp:13-1441 and a
common-name -> assembly-name alias table in src/models.hpp:1468-1706.  Those
are facts (published PhyloCSF parameters), not code.  This script re-emits
them in the reference's own *external* model file format (P.nh,
P_coding.ECM, P_noncoding.ECM; see src/ecm.hpp:21-70 and
src/models.hpp:1781-1783) so that one loader handles built-in and external
models alike.  Numeric literals are copied as text, token for token, so that
strtod() yields bit-identical doubles.

Run in the build container only (needs /root/reference):
    python tools/extract_builtin_models.py
"""
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/models.hpp")
OUT = Path(__file__).resolve().parent.parent / "phylocsfpp_b200" / "data"

CODONS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]


def tokens(body):
    return [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]


def write_ecm(path, tri, freq):
    assert len(tri) == 2016 and len(freq) == 64
    lines, k = [], 0
    for i in range(1, 64):
        lines.append(" ".join(tri[k:k + i]))
        k += i
    lines.append("")
    lines.append(" ".join(freq))
    lines.append("")
    lines.append("")
    for i in range(0, 64, 20):
        lines.append(" ".join(CODONS[i:i + 20]))
    path.write_text("\n".join(lines) + "\n")


def main():
    src = REF.read_text()
    (OUT / "models").mkdir(parents=True, exist_ok=True)
    trees = dict(re.findall(r'std::string g_(\w+)_tree = "([^"]*)";', src))
    arrays = {}
    for name, kind, body in re.findall(r"double g_(\w+?)_(cmatrix|ncmatrix|cfreq|ncfreq)\[[^\]]*\] = \{([^}]*)\};", src):
        arrays[(name, kind)] = tokens(body)
    names = re.findall(r'\{ "(\w+)", \{ &g_', src)
    assert len(names) == 11, names
    for name in names:
        (OUT / "models" / f"{name}.nh").write_text(trees[name] + "\n")
        write_ecm(OUT / "models" / f"{name}_coding.ECM", arrays[(name, "cmatrix")], arrays[(name, "cfreq")])
        write_ecm(OUT / "models" / f"{name}_noncoding.ECM", arrays[(name, "ncmatrix")], arrays[(name, "ncfreq")])
    (OUT / "models" / "INDEX").write_text("\n".join(names) + "\n")

    # alias table: { "common_name", { "asm1", "asm2" } },
    block = src[src.index("sequence_name_mapping = {"):src.index("void update_sequence_name_mapping")]
    rows = re.findall(r'\{\s*"([^"]+)",\s*\{([^}]*)\}\s*\}', block)
    with open(OUT / "species_aliases.tsv", "w") as f:
        for common, alts in rows:
            alts = re.findall(r'"([^"]*)"', alts)
            f.write(common + "\t" + ",".join(alts) + "\n")
    print(f"wrote {len(names)} models, {len(rows)} alias rows to {OUT}")


if __name__ == "__main__":
    main()
