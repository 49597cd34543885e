"""SURVEY §8 f-4 — the track consumers of the C++ host (CPU only): the bigWig reader is pinned by reproducing the reference's own
expected `annotate-with-tracks` output (test/tests.sh:23-26: example/tracks/*.bw + three example GTFs -> test/expected_results/
annotate-with-tracks/*.gtf, comment lines ignored) and the writer by an independent struct-level parser of the published format in
this file, by round trips through the reader, and by feeding its files to annotate-with-tracks in place of the reference's."""
import gzip
import os
import shutil
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phylocsfpp_b200", "bin", "phylocsf_b200")
FRAMES = ["+1", "+2", "+3", "-1", "-2", "-3"]
SETS = ["ensGene", "ncbiRefSeq", "refGene"]
CHROMS = {"chr22": 5459462, "chr25": 3980610, "chr28": 5116882}


def _need_bin():
    if not os.path.exists(BIN):
        pytest.fail(f"{BIN} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")


def _gunzip(src, dst_dir):
    dst = os.path.join(dst_dir, os.path.basename(src)[:-3])
    with gzip.open(src, "rb") as fi, open(dst, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return dst


def _body(path):
    return [ln for ln in open(path).read().split("\n") if not ln.startswith("#")]


def _expected(G, name):
    return [ln for ln in gzip.open(os.path.join(G, f"galGal6_chr22_25_28_subset_{name}.PhyloCSF++.gtf.gz"), "rt").read().split("\n") if not ln.startswith("#")]


# ---------------------------------------------------------------------------------------- an independent bigWig parser (format spec)
def parse_bigwig(path, check_zoom=True):
    b = open(path, "rb").read()
    magic, ver, nzoom, chrom_off, data_off, index_off, _fc, _dfc, _asql, summ_off, ubuf, _res = struct.unpack("<IHHQQQHHQQIQ", b[:64])
    assert magic == 0x888FFC26
    zooms = [struct.unpack("<IIQQ", b[64 + 24 * i:88 + 24 * i]) for i in range(nzoom)]
    cmagic, _bs, ks, vs, nchrom, _ = struct.unpack("<IIIIQQ", b[chrom_off:chrom_off + 32])
    assert cmagic == 0x78CA8C91 and vs == 8
    chroms = {}

    def walk_chrom(o):
        leaf, _, n = struct.unpack("<BBH", b[o:o + 4])
        for i in range(n):
            it = o + 4 + i * (ks + 8)
            key = b[it:it + ks].rstrip(b"\0").decode()
            if leaf:
                cid, clen = struct.unpack("<II", b[it + ks:it + ks + 8])
                chroms[cid] = (key, clen)
            else:
                walk_chrom(struct.unpack("<Q", b[it + ks:it + ks + 8])[0])
    walk_chrom(chrom_off + 32)
    assert len(chroms) == nchrom
    (nsec,) = struct.unpack("<Q", b[data_off:data_off + 8])
    rmagic, _blk, nitems, c0, b0, c1, b1, end_off, _ips, _ = struct.unpack("<IIQIIIIQII", b[index_off:index_off + 48])
    assert rmagic == 0x2468ACE0 and nitems == nsec
    leaves = []

    def walk_r(o):
        leaf, _, n = struct.unpack("<BBH", b[o:o + 4])
        for i in range(n):
            if leaf:
                leaves.append(struct.unpack("<IIIIQQ", b[o + 4 + 32 * i:o + 36 + 32 * i]))
            else:
                walk_r(struct.unpack("<IIIIQ", b[o + 4 + 24 * i:o + 28 + 24 * i])[4])
    walk_r(index_off + 48)
    assert len(leaves) == nsec
    intervals = []
    for (lc0, lb0, lc1, lb1, off, size) in leaves:
        d = b[off:off + size]
        if ubuf:
            d = zlib.decompress(d)
            assert len(d) <= ubuf
        cid, s, e, step, span, typ, _, cnt = struct.unpack("<IIIIIBBH", d[:24])
        assert typ == 3 and lc0 == lc1 == cid and lb0 == s and lb1 == e
        vals = np.frombuffer(d[24:24 + 4 * cnt], "<f4")
        assert s + (cnt - 1) * step + span == e
        for k, v in enumerate(vals):
            intervals.append((chroms[cid][0], s + k * step, s + k * step + span, float(v)))
    summary = struct.unpack("<Qdddd", b[summ_off:summ_off + 40])
    assert struct.unpack("<I", b[-4:])[0] == 0x888FFC26
    # zoom levels: every record inside its chromosome, ordered, and the level's valid counts add up to the bases covered
    for red, _, zdata, zindex in (zooms if check_zoom else []):          # UCSC's own files count zoom bases differently
        (nrec,) = struct.unpack("<I", b[zdata:zdata + 4])
        zmagic, _, zn = struct.unpack("<IIQ", b[zindex:zindex + 16])
        assert zmagic == 0x2468ACE0
        zl = []

        def walk_z(o):
            leaf, _, n = struct.unpack("<BBH", b[o:o + 4])
            for i in range(n):
                if leaf:
                    zl.append(struct.unpack("<IIIIQQ", b[o + 4 + 32 * i:o + 36 + 32 * i]))
                else:
                    walk_z(struct.unpack("<IIIIQ", b[o + 4 + 24 * i:o + 28 + 24 * i])[4])
        walk_z(zindex + 48)
        recs = []
        for (_a, _b, _c, _d, off, size) in zl:
            d = zlib.decompress(b[off:off + size]) if ubuf else b[off:off + size]
            recs += [struct.unpack("<IIIIffff", d[32 * i:32 * i + 32]) for i in range(len(d) // 32)]
        assert len(recs) == nrec
        assert sum(r[3] for r in recs) == summary[0]
        assert all(r[2] <= chroms[r[0]][1] and r[1] < r[2] for r in recs)
        assert all((a[0], a[2]) <= (c[0], c[1]) for a, c in zip(recs, recs[1:]))
        assert abs(sum(r[6] for r in recs) - summary[3]) <= 1e-3 * max(1.0, abs(summary[3]))
    return {"version": ver, "chroms": chroms, "intervals": intervals, "summary": summary, "zooms": zooms, "compressed": ubuf > 0}


def wig_intervals(path):
    out, chrom, pos, step, span = [], None, 0, 1, 1
    for ln in open(path):
        if ln.startswith("fixedStep"):
            kv = dict(t.split("=") for t in ln.split()[1:])
            chrom, pos, step, span = kv["chrom"], int(kv["start"]) - 1, int(kv.get("step", 1)), int(kv.get("span", 1))
        elif ln.strip():
            out.append((chrom, pos, pos + span, float(np.float32(ln))))
            pos += step
    return out


def _sizes_file(tmp):
    p = os.path.join(tmp, "chrom.sizes")
    with open(p, "w") as fh:
        for k, v in CHROMS.items():
            fh.write(f"{k}\t{v}\n")
    return p


# ---------------------------------------------------------------------------------------- tests
def test_annotate_with_tracks_reproduces_reference_expected_output(golden_dir, tmp_path):
    """test/tests.sh:23-26 — the reference's own acceptance test for this tool."""
    _need_bin()
    G = os.path.join(golden_dir, "annotate-with-tracks")
    tmp = str(tmp_path)
    gtfs = [_gunzip(os.path.join(G, f"galGal6_chr22_25_28_subset_{n}.gtf.gz"), tmp) for n in SETS]
    out = os.path.join(tmp, "annotated")
    subprocess.run([BIN, "annotate-with-tracks", "--output", out, os.path.join(G, "PhyloCSF+1.bw")] + gtfs, check=True, capture_output=True)
    for n in SETS:
        got = _body(os.path.join(out, f"galGal6_chr22_25_28_subset_{n}.PhyloCSF++.gtf"))
        assert got == _expected(G, n), n
    # default output location: next to the input; the header names the tracks
    subprocess.run([BIN, "annotate-with-tracks", os.path.join(G, "PhyloCSF+1.bw"), gtfs[2]], check=True, capture_output=True)
    side = os.path.join(tmp, "galGal6_chr22_25_28_subset_refGene.PhyloCSF++.gtf")
    assert _body(side) == _expected(G, "refGene")
    assert open(side).readline().startswith("# PhyloCSF scores computed with") and "PhyloCSF+1.bw" in open(side).readline()


def test_reader_agrees_with_independent_parser_on_reference_tracks(golden_dir):
    _need_bin()
    G = os.path.join(golden_dir, "annotate-with-tracks")
    for name in ("PhyloCSF+1.bw", "PhyloCSF-3.bw", "PhyloCSFpower.bw"):
        ref = parse_bigwig(os.path.join(G, name), check_zoom=False)
        assert {v[0]: v[1] for v in ref["chroms"].values()} == CHROMS
        dump = subprocess.run([BIN, "bigwig-dump", os.path.join(G, name)], check=True, capture_output=True, text=True).stdout.splitlines()
        rows = [ln.split("\t") for ln in dump if not ln.startswith("#")]
        assert len(rows) == len(ref["intervals"])
        for r, iv in zip(rows, ref["intervals"]):
            assert (r[0], int(r[1]), int(r[2])) == iv[:3] and np.float32(r[3]) == np.float32(iv[3])


@pytest.mark.parametrize("compress", [True, False])
def test_writer_round_trip(golden_dir, tmp_path, compress):
    """wig text -> bigWig: every interval back (independent parser and own reader), summary, zoom levels, and runs that arrive
    out of chromosome order (chr28 before chr22) are stored in tree order."""
    _need_bin()
    tmp = str(tmp_path)
    wig = _gunzip(os.path.join(golden_dir, "ref-generated", "smooth53.PhyloCSF+1.wig.gz"), tmp)
    want = wig_intervals(wig)
    # a shuffled copy: chromosomes in reverse order
    by_chrom = {}
    chrom = None
    for ln in open(wig):
        if ln.startswith("fixedStep"):
            chrom = ln.split()[1]
        by_chrom.setdefault(chrom, []).append(ln)
    shuffled = os.path.join(tmp, "shuffled.wig")
    with open(shuffled, "w") as fh:
        for c in sorted(by_chrom, reverse=True):
            fh.writelines(by_chrom[c])
    assert len(by_chrom) >= 2
    for src in (wig, shuffled):
        bw = os.path.join(tmp, "out.bw")
        subprocess.run([BIN, "wig-to-bigwig", "--compress", "1" if compress else "0", src, _sizes_file(tmp), bw], check=True, capture_output=True)
        got = parse_bigwig(bw)
        assert got["compressed"] == compress and got["version"] == 4 and len(got["zooms"]) >= 1
        assert got["intervals"] == sorted(want)
        covered, mn, mx, sm, sq = got["summary"]
        v = np.array([w[3] for w in want], np.float64)
        assert covered == 3 * len(want) and mn == v.min() and mx == v.max()
        assert abs(sm - 3 * v.sum()) < 1e-6 * abs(3 * v.sum()) and abs(sq - 3 * (v * v).sum()) < 1e-6 * 3 * (v * v).sum()
        dump = subprocess.run([BIN, "bigwig-dump", bw], check=True, capture_output=True, text=True).stdout.splitlines()
        rows = [ln.split("\t") for ln in dump if not ln.startswith("#")]
        assert [(r[0], int(r[1]), int(r[2]), float(np.float32(r[3]))) for r in rows] == sorted(want)


def test_writer_rejects_bad_input(tmp_path):
    _need_bin()
    tmp = str(tmp_path)
    sizes = _sizes_file(tmp)
    for text in ("fixedStep chrom=chrZ start=1 step=3 span=3\n1.0\n",                       # unknown chromosome
                 "fixedStep chrom=chr25 start=3980609 step=3 span=3\n1.0\n",               # runs over the chromosome end
                 "fixedStep chrom=chr22 start=10 step=3 span=3\n1\n2\nfixedStep chrom=chr22 start=12 step=3 span=3\n1\n"):  # overlap
        p = os.path.join(tmp, "bad.wig")
        open(p, "w").write(text)
        r = subprocess.run([BIN, "wig-to-bigwig", p, sizes, os.path.join(tmp, "bad.bw")], capture_output=True, text=True)
        assert r.returncode != 0 and "Error" in r.stderr


def test_annotate_from_rewritten_tracks_and_gff3(golden_dir, tmp_path):
    """The reference's tracks dumped to wig text, written back by the writer, give the same annotation as the originals (writer ->
    reader -> consumer); the example annotations are GFF3 (key=value); a GTF (key "value";) copy of a transcript gets ' key "value";' attributes; an unknown chromosome gives nan."""
    _need_bin()
    G = os.path.join(golden_dir, "annotate-with-tracks")
    tmp = str(tmp_path)
    sizes = _sizes_file(tmp)
    tdir = os.path.join(tmp, "tracks")
    os.makedirs(tdir)
    for name in [f"PhyloCSF{f}" for f in FRAMES] + ["PhyloCSFpower"]:
        dump = subprocess.run([BIN, "bigwig-dump", os.path.join(G, name + ".bw")], check=True, capture_output=True, text=True).stdout.splitlines()
        wig = os.path.join(tdir, name + ".wig")
        with open(wig, "w") as fh:
            prev = None
            for ln in dump:
                if ln.startswith("#"):
                    continue
                c, b, e, v = ln.split("\t")
                if prev != (c, int(b)):
                    fh.write(f"fixedStep chrom={c} start={int(b) + 1} step={int(e) - int(b)} span={int(e) - int(b)}\n")
                fh.write(v + "\n")
                prev = (c, int(e))
        subprocess.run([BIN, "wig-to-bigwig", wig, sizes, os.path.join(tdir, name + ".bw")], check=True, capture_output=True)
    gtf = _gunzip(os.path.join(G, "galGal6_chr22_25_28_subset_ensGene.gtf.gz"), tmp)
    out = os.path.join(tmp, "o")
    subprocess.run([BIN, "annotate-with-tracks", "--output", out, os.path.join(tdir, "PhyloCSF+1.bw"), gtf], check=True, capture_output=True)
    assert _body(os.path.join(out, "galGal6_chr22_25_28_subset_ensGene.PhyloCSF++.gtf")) == _expected(G, "ensGene")
    # a .wig path gets the reference's hint
    r = subprocess.run([BIN, "annotate-with-tracks", os.path.join(tdir, "PhyloCSF+1.wig"), gtf], capture_output=True, text=True)
    assert r.returncode != 0 and "wigToBigWig" in r.stdout
    # GTF attributes + unknown chromosome
    gff = os.path.join(tmp, "x.gtf")
    lines = [ln for ln in open(gtf).read().split("\n") if ln and not ln.startswith("#")]
    first = []
    for ln in lines:          # the first transcript with CDS lines
        cols = ln.split("\t")
        if cols[2] == "transcript" and first:
            if any(x.split("\t")[2] == "CDS" for x in first):
                break
            first = []
        first.append(ln)
    with open(gff, "w") as fh:
        for ln in first:
            cols = ln.split("\t")
            cols[8] = 'gene_id "x"; transcript_id "y";'
            fh.write("\t".join(cols) + "\n")
        for ln in first:
            cols = ln.split("\t")
            cols[0] = "chrUn"
            cols[8] = 'gene_id "x"; transcript_id "y";'
            fh.write("\t".join(cols) + "\n")
    r = subprocess.run([BIN, "annotate-with-tracks", "--output", out, os.path.join(G, "PhyloCSF+1.bw"), gff], check=True, capture_output=True, text=True)
    assert "chrUn" in r.stdout
    got = _body(os.path.join(out, "x.PhyloCSF++.gtf"))
    want = {}
    for ln in _expected(G, "ensGene"):
        if "phylocsf_score_weighted_mean" in ln:
            want.setdefault(tuple(ln.split("\t")[:8]), (ln.split("phylocsf_score_weighted_mean=")[1].split(";")[0], ln.split("phylocsf_power_mean=")[1]))
    n_checked = 0
    for g, f in zip(got[:len(first)], first):
        if f.split("\t")[2] in ("transcript", "CDS"):
            score, power = want[tuple(f.split("\t")[:8])]
            assert g.endswith(f'gene_id "x"; transcript_id "y"; phylocsf_score_weighted_mean "{score}"; phylocsf_power_mean "{power}";'), g
            n_checked += 1
        else:
            assert g.endswith('gene_id "x"; transcript_id "y";')
    assert n_checked >= 2
    for g, f in zip(got[len(first):2 * len(first)], first):
        if f.split("\t")[2] in ("transcript", "CDS"):
            assert g.endswith(' phylocsf_score_weighted_mean "nan"; phylocsf_power_mean "nan";')
