"""CPU tests of the C++ host (libnimpress_host.so): parsers, VCF/BCF readers, the streaming
findVariant, stats and output formatting -- against the oracle and the reference's known answers.
No GPU needed: nothing here computes a score."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import orc
from util_bcf import write_bcf, write_vcf
from util_files import make_dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
S1 = os.path.join(G, "set1")


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    api.load_host_library()
    return api


def plan(api, score, vcf, bed=None, ignorefilt=False):
    L = api.load_host_library()
    p = api._Params(0, 0, 3, int(ignorefilt), int(bed is not None), 0, 0, 0, 100, 0.05, 0.001)
    kind = np.zeros(1 << 16, np.int32); ea = np.zeros(1 << 16, np.int32)
    n_rows, n_s = C.c_int64(), C.c_int64()
    rc = L.nph_plan(os.fsencode(score), os.fsencode(vcf), os.fsencode(bed) if bed else None, C.byref(p),
                    kind.ctypes.data, ea.ctypes.data, len(kind), C.byref(n_rows), C.byref(n_s))
    return rc, kind[:n_rows.value].copy(), ea[:n_rows.value].copy(), n_s.value


def read_gt(api, path, row_bytes, max_records=4096):
    L = api.load_host_library()
    out = np.zeros((max_records, row_bytes), np.uint8)
    nrec, ns, w, pl = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    rc = L.nph_read_gt(os.fsencode(path), out.ctypes.data, row_bytes, max_records, C.byref(nrec), C.byref(ns), C.byref(w), C.byref(pl))
    assert rc == 0, rc
    return out[:nrec.value], ns.value, w.value, pl.value


def oracle_plan(score, vcf, bed=None, ignorefilt=False):
    r = orc.compute_scores_files(score, vcf, bed, maxmis=1.0, ignorefilt=ignorefilt, skip_aftest=True)
    kind = np.where(r["loci"]["klass"] == orc.CLASS_MAXMIS, 0, r["loci"]["klass"])
    return kind, r["loci"]["eaidx"]


def test_exports_match_header(api):
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nimpress_host.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(nph_[a-z_0-9]+)\s*\(", txt)))
    L = api.load_host_library()
    for s in syms:
        assert hasattr(L, s), s
    assert syms == L._nph_symbols


def test_stats_known_answers(api):
    """tests/test_stats.nim through the product's stats (bisection instead of enumeration)."""
    L = api.load_host_library()
    fns = {"dbinom": L.nph_dbinom, "pbinom": L.nph_pbinom, "binom_test": L.nph_binom_test, "betai": L.nph_betai}
    for k in json.load(open(os.path.join(G, "stats_kat.json")))["kats"]:
        a = k["args"]
        v = fns[k["fn"]](a[0], a[1], a[2]) if k["fn"] == "betai" else fns[k["fn"]](int(a[0]), int(a[1]), a[2])
        t = k["target"]
        if k["kind"] == "exact":
            assert v == t, k
        elif abs(t) < 1e-9:
            assert abs(v - t) < 1e-9, k
        else:
            assert abs((v - t) / t) < 1e-5, k


def test_binom_test_bisection_equals_enumeration(api):
    """The O(log n) tail search returns the same p-value as the reference's O(n) enumeration."""
    L, O = api.load_host_library(), orc.lib()
    rng = np.random.default_rng(5)
    for n in (2, 7, 40, 1000, 20000, 200000):
        for _ in range(60):
            p = float(rng.choice([rng.uniform(0.001, 0.999), 0.5, 0.25, 0.01]))
            x = int(rng.integers(0, n + 1)) if rng.random() < 0.5 else int(np.clip(rng.normal(n * p, 3 * np.sqrt(n * p * (1 - p)) + 1), 0, n))
            a, b = L.nph_binom_test(x, n, p), O.orc_binom_test(x, n, p)
            assert a == b or abs(a - b) <= 1e-12 * max(abs(b), 1e-300), (x, n, p, a, b)


def test_float_format_corpus(api):
    L = api.load_host_library()
    buf = C.create_string_buffer(64)
    for v in open(os.path.join(G, "res_format_corpus.tsv")).read().split():
        L.nph_format_float(float(v), buf, 64)
        assert buf.value.decode() == v
    for x, s in ((2.0, "2.0"), (float("nan"), "nan"), (-float("inf"), "-inf"), (1e22, "1e+22"), (-3.0, "-3.0"), (0.1, "0.1")):
        L.nph_format_float(x, buf, 64)
        assert buf.value.decode() == s


@pytest.mark.parametrize("bed,ign", [(None, False), (os.path.join(S1, "set1.bed"), False), (None, True)])
def test_plan_set1(api, bed, ign):
    rc, kind, ea, n = plan(api, os.path.join(S1, "set1.score"), os.path.join(S1, "set1.vcf.gz"), bed, ign)
    assert rc == 0 and n == 6
    ok, oea = oracle_plan(os.path.join(S1, "set1.score"), os.path.join(S1, "set1.vcf.gz"), bed, ign)
    assert list(kind) == list(ok) and list(ea) == list(oea)


def test_set1_gt_payload(api):
    """CRLF BGZF VCF text -> int8 diploid payload identical to a hand transcription of the file."""
    gt, n, w, pl = read_gt(api, os.path.join(S1, "set1.vcf.gz"), 16)
    assert (n, w, pl, len(gt)) == (6, 1, 2, 7)
    # record 5: 1:300 GA>T,CT   0/0 2/2 0/1 1/0 ./. 1/1
    assert list(gt[4, :12].view(np.int8)) == [2, 2, 6, 6, 2, 4, 4, 2, 0, 0, 4, 4]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_plan_and_payload_vcf_bcf_oracle(api, tmp_path, seed):
    """Streaming findVariant on VCF text and on BCF == the oracle's scan, incl. duplicate sites,
    multi-base REF overlap at another POS, INFO/END spans, missing ALT, unknown contigs, and the
    BED off-by-one traps; both readers hand out byte-identical GT payloads."""
    rng = np.random.default_rng(seed)
    d = make_dataset(str(tmp_path), rng, n=50, V=90, sorted_scores=seed != 3)
    for bed in (None, d["bed"]):
        for ign in (False, True):
            ok, oea = oracle_plan(d["score"], d["vcf"], bed, ign)
            for f in (d["vcf"], d["bcf"]):
                rc, kind, ea, n = plan(api, d["score"], f, bed, ign)
                assert rc == 0 and n == 50
                assert np.array_equal(kind, ok), f
                assert np.array_equal(ea, oea), f
    a, n1, w1, p1 = read_gt(api, d["vcf"], 128)
    b, n2, w2, p2 = read_gt(api, d["bcf"], 128)
    assert (n1, w1, p1) == (n2, w2, p2) == (50, 1, 2) and np.array_equal(a, b)
    want = np.stack([r["gt"].reshape(-1).view(np.uint8) for r in d["records"]])
    assert np.array_equal(a[:, :100], want)


def test_bcf_variants(api, tmp_path):
    """BCF details: IDX= permuted dictionary, a FORMAT field before GT, int16 GT, haploid calls
    padded with vector_end, uncompressed BCF, gzip (non-BGZF) VCF."""
    rng = np.random.default_rng(9)
    d = make_dataset(str(tmp_path), rng, n=20, V=30, haploid_rate=0.2)
    ok, oea = oracle_plan(d["score"], d["vcf"], None, False)
    contigs = ["1", "2", "X"]
    for name, kw in (("idx.bcf", dict(with_idx=True)), ("fmt.bcf", dict(extra_fmt=True)), ("raw.bcf", dict(compress=None))):
        p = os.path.join(str(tmp_path), name)
        write_bcf(p, d["samples"], d["records"], contigs, filters=("FAIL", "LowQ"), **kw)
        rc, kind, ea, n = plan(api, d["score"], p)
        assert rc == 0 and np.array_equal(kind, ok) and np.array_equal(ea, oea), name
        a, *_ = read_gt(api, p, 64)
        b, *_ = read_gt(api, d["bcf"], 64)
        assert np.array_equal(a, b), name
    p = os.path.join(str(tmp_path), "gz.vcf.gz")
    write_vcf(p, d["samples"], d["records"], contigs=contigs, filters=("FAIL", "LowQ"), compress="gzip", crlf=True)
    rc, kind, ea, n = plan(api, d["score"], p)
    assert rc == 0 and np.array_equal(kind, ok)
    recs16 = [dict(r, gt=r["gt"].astype(np.int16)) for r in d["records"]]
    for r in recs16:
        r["gt"][r["gt"] == -127] = -32767
    p = os.path.join(str(tmp_path), "w16.bcf")
    write_bcf(p, d["samples"], recs16, contigs, filters=("FAIL", "LowQ"))
    a, n, w, pl = read_gt(api, p, 128)
    assert (w, pl) == (2, 2) and np.array_equal(a[:, :80].view(np.int16), np.stack([r["gt"].reshape(-1) for r in recs16]))


def test_input_errors(api, tmp_path):
    """The reference's doAssert / ValueError paths surface as NPH_EINPUT, unopenable files as -1/-2."""
    vcf = os.path.join(S1, "set1.vcf.gz")
    good = open(os.path.join(S1, "set1.score")).read()
    cases = {"blank_tail": good + "\n\n", "five_cols": good + "\n1\t2\tA\tC\t0.1", "bad_float": good.replace("0.123", "abc"),
             "bad_pos": good.replace("\n1\t100\t", "\n1\tx100\t")}
    for name, txt in cases.items():
        p = os.path.join(str(tmp_path), name)
        open(p, "w").write(txt)
        assert plan(api, p, vcf)[0] == -3, name
    assert plan(api, os.path.join(str(tmp_path), "missing"), vcf)[0] == -2
    assert plan(api, os.path.join(S1, "set1.score"), os.path.join(str(tmp_path), "missing.vcf"))[0] == -1
    bed = os.path.join(str(tmp_path), "two_cols.bed")
    open(bed, "w").write("1\t5\n")
    assert plan(api, os.path.join(S1, "set1.score"), vcf, bed)[0] == -3
    notvcf = os.path.join(str(tmp_path), "not.vcf")
    open(notvcf, "w").write("hello\n")
    assert plan(api, os.path.join(S1, "set1.score"), notvcf)[0] == -1


def test_bgzf_thread_pool_equals_sequential(api, tmp_path, monkeypatch):
    """Multi-block BGZF files (dozens of 64 KiB blocks) through the worker-pool inflater give the same
    bytes as the sequential zlib stream, for BCF payloads and for VCF text."""
    rng = np.random.default_rng(4)
    d = make_dataset(str(tmp_path), rng, n=3000, V=240, with_traps=False)
    assert os.path.getsize(d["bcf"]) > 300_000
    outs = {}
    for th in ("1", "2", "7"):
        monkeypatch.setenv("NIMPRESS_THREADS", th)
        a, n1, w1, p1 = read_gt(api, d["bcf"], 6016, max_records=400)
        b, n2, w2, p2 = read_gt(api, d["vcf"], 6016, max_records=400)
        assert np.array_equal(a, b) and len(a) == len(d["records"])
        outs[th] = a
        rc, kind, ea, n = plan(api, d["score"], d["bcf"])
        assert rc == 0 and n == 3000
    assert np.array_equal(outs["1"], outs["2"]) and np.array_equal(outs["1"], outs["7"])
    want = np.stack([r["gt"].reshape(-1).view(np.uint8) for r in d["records"]])
    assert np.array_equal(outs["7"][:, :6000], want)
    # a truncated file is an input error, not a hang
    bad = os.path.join(str(tmp_path), "trunc.bcf")
    open(bad, "wb").write(open(d["bcf"], "rb").read()[:200_000])
    monkeypatch.setenv("NIMPRESS_THREADS", "4")
    L = api.load_host_library()
    out = np.zeros((400, 6016), np.uint8)
    nrec, ns, w, pl = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    assert L.nph_read_gt(os.fsencode(bad), out.ctypes.data, 6016, 400, C.byref(nrec), C.byref(ns), C.byref(w), C.byref(pl)) == -3


def test_vcf_text_gt_paths(api, tmp_path):
    """Hand-written VCF text lines through both GT parsers of the text reader: the direct int8 path
    (diploid, one-character alleles, GT first) and the general one (other ploidies, allele numbers >= 10,
    GT not first, short lines), each expected byte for byte: (allele+1)<<1|phased, 0|phase = missing,
    vector_end (0x81 / 0x8001) padding."""
    hdr = "##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\td\n"
    lines = [
        "1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0/0\t0|1\t./.\t1/.",                 # direct
        "1\t20\t.\tA\tC,G,T\t.\tPASS\t.\tGT:DP\t3/0:7\t.|2:1\t9|9:0\t0/1:.",  # direct, sub-fields after GT
        "1\t30\t.\tA\tC\t.\tPASS\t.\tGT\t0/1\t1\t0/1/1\t.",                   # general: haploid and triploid calls
        "1\t40\t.\tA\tC\t.\tPASS\t.\tDP:GT\t5:0/1\t6:1|1\t7:./.\t8:0|0",      # general: GT second
        "1\t50\t.\tA\tC\t.\tPASS\t.\tGT\t0/1\t1/1",                           # general: two sample columns missing
        "1\t60\t.\tA\t" + ",".join("C" * (i + 2) for i in range(70)) + "\t.\tPASS\t.\tGT\t0/70\t12|3\t./10\t0/0",   # general: int16
        "1\t70\t.\tA\tC\t.\tPASS\t.\tDP\t1\t2\t3\t4",                         # no GT at all
    ]
    p = tmp_path / "t.vcf"
    p.write_text(hdr + "\n".join(lines) + "\n")
    L = api.load_host_library()
    out = np.zeros((8, 64), np.uint8)
    nrec, ns, w, pl = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    # nph_read_gt reports the layout of the last GT-bearing record; rows are checked per record below
    assert L.nph_read_gt(os.fsencode(str(p)), out.ctypes.data, 64, 8, C.byref(nrec), C.byref(ns), C.byref(w), C.byref(pl)) == 0
    assert (nrec.value, ns.value) == (7, 4)
    i8 = lambda row, k: list(out[row, :k].view(np.int8))
    assert i8(0, 8) == [2, 2, 2, 5, 0, 0, 4, 0]
    assert i8(1, 8) == [8, 2, 0, 7, 20, 21, 2, 4]
    VE = -127
    assert i8(2, 12) == [2, 4, VE, 4, VE, VE, 2, 4, 4, 0, VE, VE]
    assert i8(3, 8) == [2, 4, 4, 5, 0, 0, 2, 3]
    assert i8(4, 8) == [2, 4, 4, 4, 0, VE, 0, VE]
    assert list(out[5, :16].view(np.int16)) == [2, 142, 26, 9, 0, 22, 2, 2]
    assert not out[6].any()


@pytest.mark.parametrize("seed", range(5))
def test_vcf_text_gt_random_grammar(api, tmp_path, seed):
    """Random GT strings (ploidy 1-3, '/' and '|', '.', allele numbers up to 150, optional sub-fields after
    GT, GT sometimes not first) against an encoder written here from the BCF rules: which of the text
    reader's two parsers a line takes must not be observable."""
    rng = np.random.default_rng(seed)
    n = 37
    hdr = "##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n)) + "\n"
    lines, want = [], []
    for rec in range(40):
        style = rng.integers(4)                     # 0: plain diploid small alleles (direct path), 1: + sub-fields, 2: mixed ploidy / big alleles, 3: GT second
        max_allele = 150 if (style == 2 and rng.random() < 0.3) else (12 if style == 2 else 9)
        cols, vals = [], []
        for s in range(n):
            pl = 2 if style in (0, 1) else int(rng.integers(1, 4))
            toks, enc = [], []
            for k in range(pl):
                miss = rng.random() < 0.1
                a = int(rng.integers(0, max_allele + 1))
                sep = "|" if (k and rng.random() < 0.4) else "/"
                toks.append((sep if k else "") + ("." if miss else str(a)))
                enc.append((0 if miss else (a + 1) << 1) | (1 if (k and sep == "|") else 0))
            gt = "".join(toks)
            if style == 1:
                gt += ":" + str(int(rng.integers(0, 99)))
            if style == 3:
                gt = str(int(rng.integers(0, 99))) + ":" + gt
            cols.append(gt); vals.append(enc)
        fmt = {0: "GT", 1: "GT:DP", 2: "GT", 3: "DP:GT"}[int(style)]
        lines.append(f"1\t{100 + rec}\t.\tA\t" + ",".join("C" * (i + 2) for i in range(max_allele)) + f"\t.\tPASS\t.\t{fmt}\t" + "\t".join(cols))
        want.append(vals)
    p = tmp_path / "r.vcf"
    p.write_text(hdr + "\n".join(lines) + "\n")
    L = api.load_host_library()
    row_bytes = n * 3 * 2
    # one record at a time would need an API; instead read all rows and decode each with its own layout
    out = np.zeros((len(lines), row_bytes), np.uint8)
    nrec, ns, w, pl = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    assert L.nph_read_gt(os.fsencode(str(p)), out.ctypes.data, row_bytes, len(lines), C.byref(nrec), C.byref(ns), C.byref(w), C.byref(pl)) == 0
    assert (nrec.value, ns.value) == (len(lines), n)
    for r, vals in enumerate(want):
        ploidy = max(len(v) for v in vals)
        width = 1 if max(max(v) for v in vals) <= 127 else 2
        dt, vend = (np.int8, -127) if width == 1 else (np.int16, -32767)
        exp = np.full((n, ploidy), vend, dtype=dt)
        for s, v in enumerate(vals):
            exp[s, :len(v)] = v
        got = out[r, :n * ploidy * width].view(dt).reshape(n, ploidy)
        assert np.array_equal(got, exp), (r, lines[r][:80])


def plan_indexed(api, score, vcf, bed=None, ignorefilt=False):
    L = api.load_host_library()
    p = api._Params(0, 0, 3, int(ignorefilt), int(bed is not None), 0, 0, 0, 100, 0.05, 0.001)
    kind = np.zeros(1 << 16, np.int32); ea = np.zeros(1 << 16, np.int32)
    n_rows, n_s, nrec, seeks = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.nph_plan_indexed(os.fsencode(score), os.fsencode(vcf), os.fsencode(bed) if bed else None, C.byref(p), kind.ctypes.data,
                            ea.ctypes.data, len(kind), C.byref(n_rows), C.byref(n_s), C.byref(nrec), C.byref(seeks))
    assert rc == 0, rc
    return kind[:n_rows.value].copy(), ea[:n_rows.value].copy(), nrec.value, seeks.value


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_index_driven_reader_equals_streaming(api, tmp_path, seed, monkeypatch):
    """With <file>.tbi (VCF) / <file>.csi (BCF) the reader jumps between the score's loci like the
    reference's per-locus index queries; what each score row matches -- duplicate sites, a 2-base REF
    overlapping the next position, INFO/END spans reaching across index windows, contigs missing from
    the file -- is identical to streaming the whole file, on far fewer records."""
    import util_bcf
    rng = np.random.default_rng(seed)
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 3000 + 500 * seed)           # ~40 BGZF blocks: jumps cross block boundaries
    d = make_dataset(str(tmp_path), rng, n=12, V=1500, index=True, spread=400, sorted_scores=seed % 2 == 0)
    # a sparse score: a few of the dataset's entries (incl. every trap entry at the end), so that the index pays off
    offset_and_entries = open(d["score"]).read().split("\n")
    head, ents = offset_and_entries[:5], offset_and_entries[5:]
    keep = [e for i, e in enumerate(ents) if e and (rng.random() < 0.03 or i >= len(ents) - 12)]
    sparse = tmp_path / "sparse.score"
    sparse.write_text("\n".join(head + keep) + "\n")
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    for f in (d["vcf"], d["bcf"]):
        for bed in (None, d["bed"]):
            rc, kind, ea, n = plan(api, str(sparse), f, bed)
            k2, e2, nrec, seeks = plan_indexed(api, str(sparse), f, bed)
            assert rc == 0 and np.array_equal(kind, k2) and np.array_equal(ea, e2), f
            assert seeks > 0 and nrec < len(d["records"]) // 2, (f, seeks, nrec)
    # a dense score reads everything: same answers again
    k1 = plan(api, d["score"], d["bcf"])
    k2 = plan_indexed(api, d["score"], d["bcf"])
    assert np.array_equal(k1[1], k2[0]) and np.array_equal(k1[2], k2[1])
    # without the override the library weighs regions against file size; with NIMPRESS_NO_INDEX it never uses the index
    monkeypatch.delenv("NIMPRESS_FORCE_INDEX")
    monkeypatch.setenv("NIMPRESS_NO_INDEX", "1")
    assert plan_indexed(api, str(sparse), d["bcf"])[3] == 0


def test_index_reader_long_spans_and_edges(api, tmp_path, monkeypatch):
    """Records whose INFO/END reaches over many 16 kb index windows (they sit in a coarse bin, far before the
    locus they overlap), loci before the first / after the last record, on a contig the file lacks, and at the
    first base: the index-driven pass finds exactly what streaming finds (TBI and CSI)."""
    import util_bcf
    from util_bcf import write_bcf, write_vcf
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 2000)
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    rng = np.random.default_rng(8)
    n = 6
    samples = [f"s{i}" for i in range(n)]
    g = lambda: ((rng.integers(0, 2, size=(n, 2)) + 1) << 1).astype(np.int8)
    recs = []
    for c in ("1", "2"):
        recs.append(dict(contig=c, pos=5000, ref="G", alts=["T"], filter="PASS", info="END=405000", gt=g()))     # spans 25 windows
        for p in range(20000, 900000, 1700):
            recs.append(dict(contig=c, pos=p, ref="A", alts=["C"], filter="PASS", gt=g()))
        recs.append(dict(contig=c, pos=300000, ref="AT", alts=["A"], filter="PASS", gt=g()))
    order = {"1": 0, "2": 1}
    recs.sort(key=lambda r: (order[r["contig"]], r["pos"]))
    vcf, bcf = str(tmp_path / "l.vcf.gz"), str(tmp_path / "l.bcf")
    write_vcf(vcf, samples, recs, contigs=["1", "2"], compress="bgzf", index="tbi")
    write_bcf(bcf, samples, recs, ["1", "2"], index=True)
    ents = [("1", 1, "G", "T"), ("1", 4999, "G", "T"), ("1", 5000, "G", "T"), ("1", 200000, "G", "T"), ("1", 404999, "G", "G"), ("1", 405001, "G", "T"),
            ("1", 300001, "AT", "AT"), ("2", 333300, "G", "T"), ("2", 899700, "A", "C"), ("2", 2000000, "A", "C"), ("3", 100, "A", "C"),
            ("1", 21700, "A", "C"), ("2", 20000, "A", "A")]
    sc = tmp_path / "l.score"
    sc.write_text("x\nd\nc\nhs37d5\n0\n" + "\n".join(f"{c}\t{p}\t{r}\t{e}\t0.1\t0.2" for c, p, r, e in ents) + "\n")
    for f in (vcf, bcf):
        rc, kind, ea, _ = plan(api, str(sc), f)
        k2, e2, nrec, seeks = plan_indexed(api, str(sc), f)
        assert rc == 0 and np.array_equal(kind, k2) and np.array_equal(ea, e2), (f, kind, k2)
        assert seeks > 0 and nrec < len(recs) // 3
        assert list(kind[:7]) == [2, 2, 0, 0, 0, 2, 0] and kind[9] == 2 and kind[10] == 2      # 0 = matched, 2 = absent


def test_fast_inflate_equals_zlib(api):
    """The reader's own DEFLATE decoder against zlib: stored, fixed and dynamic blocks, every zlib strategy, run-heavy /
    random / periodic data, lengths around the word-copy edges; then flipped bits -- whatever it accepts wrongly is what
    the CRC32 check behind it is for (it must never crash or write out of bounds)."""
    import zlib
    L = api.load_host_library()
    rng = np.random.default_rng(1)

    def run(comp, n):
        inb = np.frombuffer(comp + b"\0" * 16, dtype=np.uint8).copy()
        out = np.full(n + 48, 0xAA, np.uint8)
        ok = L.nph_fast_inflate(inb.ctypes.data, len(comp), out.ctypes.data, n)
        assert np.all(out[n + 16:] == 0xAA)                      # never beyond the documented slack (16 bytes)
        return ok, out[:n].tobytes()

    for n in (0, 1, 2, 7, 8, 9, 15, 16, 17, 100, 1000, 65280):
        datas = [bytes(rng.integers(0, 256, n, dtype=np.uint8)), bytes(rng.integers(0, 3, n, dtype=np.uint8) * 2),
                 (b"\x02\x02\x02\x04" * (n // 4 + 1))[:n], bytes(np.repeat(rng.integers(0, 256, max(n // 50, 1), dtype=np.uint8), 50)[:n]),
                 b"a" * n, bytes((np.arange(n) % 251).astype(np.uint8)), (b"abcdefg" * (n // 7 + 1))[:n]]
        for d in datas:
            for level in (0, 1, 6, 9):
                for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strat)
                    comp = co.compress(d) + co.flush()
                    ok, got = run(comp, len(d))
                    assert ok == 1 and got == d, (n, level, strat)
                    assert run(comp, len(d) + 1)[0] == 0 and (len(d) == 0 or run(comp, len(d) - 1)[0] == 0)   # wrong expected size
    d = bytes(rng.integers(0, 4, 30000, dtype=np.uint8))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(d) + co.flush()
    for trial in range(400):
        c = bytearray(comp)
        i = int(rng.integers(0, len(c)))
        c[i] ^= 1 << int(rng.integers(0, 8))
        run(bytes(c), len(d))
    assert run(comp[:len(comp) // 2], len(d))[0] == 0               # truncated


def test_bgzf_pool_decoders_agree_and_crc_is_checked(api, tmp_path, monkeypatch):
    """Pool with this engine's decoder == pool with zlib only == sequential zlib; a flipped payload bit in a block is
    reported (CRC32), not scored."""
    rng = np.random.default_rng(5)
    d = make_dataset(str(tmp_path), rng, n=300, V=400)
    monkeypatch.setenv("NIMPRESS_THREADS", "4")
    a = read_gt(api, d["bcf"], 1024)
    monkeypatch.setenv("NIMPRESS_ZLIB_ONLY", "1")
    b = read_gt(api, d["bcf"], 1024)
    monkeypatch.setenv("NIMPRESS_THREADS", "1")
    c = read_gt(api, d["bcf"], 1024)
    assert a[1:] == b[1:] == c[1:] and np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])
    monkeypatch.delenv("NIMPRESS_ZLIB_ONLY")
    monkeypatch.setenv("NIMPRESS_THREADS", "4")
    raw = bytearray(open(d["bcf"], "rb").read())
    raw[len(raw) // 2] ^= 0x10
    bad = tmp_path / "bad.bcf"
    bad.write_bytes(bytes(raw))
    L = api.load_host_library()
    out = np.zeros((4096, 1024), np.uint8)
    nrec, ns, w, pl = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    rc = L.nph_read_gt(os.fsencode(str(bad)), out.ctypes.data, 1024, 4096, C.byref(nrec), C.byref(ns), C.byref(w), C.byref(pl))
    assert rc == -3 and b"BGZF" in L.nph_last_error()


def test_broken_index_files_are_ignored(api, tmp_path, monkeypatch):
    """A truncated, empty or garbage .tbi / .csi must not crash the reader or change what is matched: it does not
    parse, so the file is streamed."""
    import util_bcf
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 4000)
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    rng = np.random.default_rng(12)
    d = make_dataset(str(tmp_path), rng, n=8, V=300, index=True, spread=300)
    want = plan(api, d["score"], d["bcf"])
    for f, ext in ((d["vcf"], ".tbi"), (d["bcf"], ".csi")):
        import gzip
        raw = gzip.decompress(open(f + ext, "rb").read())
        for cut in (0, 3, 4, 11, 40, len(raw) // 2, len(raw) - 1):
            open(f + ext, "wb").write(util_bcf.bgzf_compress(raw[:cut]))
            k, e, nrec, seeks = plan_indexed(api, d["score"], f)
            assert seeks == 0 and nrec == len(d["records"]) and np.array_equal(k, want[1]) and np.array_equal(e, want[2]), (ext, cut)
        open(f + ext, "wb").write(bytes(rng.integers(0, 256, 500, dtype=np.uint8)))          # not even gzip
        k, e, nrec, seeks = plan_indexed(api, d["score"], f)
        assert seeks == 0 and np.array_equal(k, want[1])
        open(f + ext, "wb").write(util_bcf.bgzf_compress(raw))                               # intact again
        k, e, _, _ = plan_indexed(api, d["score"], f)
        assert np.array_equal(k, want[1]) and np.array_equal(e, want[2])


def test_stale_index_is_not_used(api, tmp_path, monkeypatch):
    """An index older than its data file (the file was rewritten since) is ignored: the file is streamed."""
    import util_bcf
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 4000)
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    rng = np.random.default_rng(13)
    d = make_dataset(str(tmp_path), rng, n=8, V=600, index=True, spread=300)
    lines = open(d["score"]).read().split("\n")
    sparse = tmp_path / "s.score"
    sparse.write_text("\n".join(lines[:5] + [e for i, e in enumerate(lines[5:]) if e and i % 40 == 0]) + "\n")
    assert plan_indexed(api, str(sparse), d["bcf"])[3] > 0
    st = os.stat(d["bcf"] + ".csi")
    os.utime(d["bcf"], (st.st_atime + 100, st.st_mtime + 100))           # data "modified" after the index was made
    k, e, nrec, seeks = plan_indexed(api, str(sparse), d["bcf"])
    assert seeks == 0 and nrec == len(d["records"])


def test_index_with_contig_blocks_out_of_header_order(api, tmp_path, monkeypatch):
    """htslib only requires each contig's records to be contiguous: a file may hold its contig blocks in another
    order than its header lists them, and index fine.  The index-driven pass walks contigs in header order and
    only seeks forward, so such a file is streamed instead (round-1 advisor finding: its loci on the later-named,
    earlier-stored contig came back ABSENT)."""
    import util_bcf
    from util_bcf import write_bcf, write_vcf
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 1500)
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    rng = np.random.default_rng(21)
    n = 5
    samples = [f"s{i}" for i in range(n)]
    g = lambda: ((rng.integers(0, 2, size=(n, 2)) + 1) << 1).astype(np.int8)
    recs = []
    for c in ("chr2", "chr1"):                                 # file order: chr2's block first
        for p in range(1000, 400000, 900):
            recs.append(dict(contig=c, pos=p, ref="A", alts=["C"], filter="PASS", gt=g()))
    vcf, bcf = str(tmp_path / "p.vcf.gz"), str(tmp_path / "p.bcf")
    write_vcf(vcf, samples, recs, contigs=["chr1", "chr2"], compress="bgzf", index="tbi")
    write_bcf(bcf, samples, recs, ["chr1", "chr2"], index=True)
    ents = [("chr1", 1000, "A", "C"), ("chr1", 300700, "A", "C"), ("chr2", 1900, "A", "C"), ("chr2", 399700, "A", "A"), ("chr2", 5, "A", "C")]
    sc = tmp_path / "p.score"
    sc.write_text("x\nd\nc\nhs37d5\n0\n" + "\n".join(f"{c}\t{p}\t{r}\t{e}\t0.1\t0.2" for c, p, r, e in ents) + "\n")
    for f in (vcf, bcf):
        rc, kind, ea, _ = plan(api, str(sc), f)
        k2, e2, nrec, seeks = plan_indexed(api, str(sc), f)
        assert rc == 0 and list(kind) == [0, 0, 0, 0, 2], (f, kind)
        assert np.array_equal(kind, k2) and np.array_equal(ea, e2), (f, kind, k2)
    # one contig only: the index is still used
    sc1 = tmp_path / "p1.score"
    sc1.write_text("x\nd\nc\nhs37d5\n0\n" + "\n".join(f"{c}\t{p}\t{r}\t{e}\t0.1\t0.2" for c, p, r, e in ents[:2]) + "\n")
    k2, e2, nrec, seeks = plan_indexed(api, str(sc1), bcf)
    assert list(k2) == [0, 0] and seeks > 0


def test_malformed_bgzf_extra_field_is_rejected(api, tmp_path):
    """A BGZF header claiming XLEN < 6 while carrying the BC subfield (round-1 advisor finding: the pool's reader
    underflowed `xlen - 6` and read before its buffer): the file must be refused, not read."""
    import struct
    import zlib
    payload = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\n1\t5\t.\tA\tC\t.\tPASS\t.\tGT\t0/1\n"
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(payload) + co.flush()
    for xlen in (0, 2, 5):
        bsize = 18 + len(body) + 8 - 1
        hdr = b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<H", xlen) + b"BC" + struct.pack("<HH", 2, bsize)
        blk = hdr + body + struct.pack("<II", zlib.crc32(payload), len(payload))
        f = tmp_path / f"bad{xlen}.vcf.gz"
        f.write_bytes(blk + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        sc = tmp_path / "s.score"
        sc.write_text("x\nd\nc\nhs37d5\n0\n1\t5\tA\tC\t0.1\t0.2\n")
        kind = np.zeros(4, np.int32); ea = np.zeros(4, np.int32)
        nr, ns = C.c_int64(), C.c_int64()
        L = api.load_host_library()
        p = api._Params(0, 0, 3, 0, 0, 0, 0, 0, 100, 0.05, 0.001)
        rc = L.nph_plan(os.fsencode(str(sc)), os.fsencode(str(f)), None, C.byref(p), kind.ctypes.data, ea.ctypes.data, 4, C.byref(nr), C.byref(ns))
        assert rc != 0, xlen                                  # an error code, not a crash and not a silent read


def test_score_fields_are_parsed_as_strictly_as_nim(api, tmp_path):
    """strtod / strtoll accept what Nim's parseFloat / parseInt raise ValueError on: leading blanks, hex floats,
    nan(...).  Such score files must fail like the reference instead of scoring."""
    d = tmp_path
    (d / "v.vcf").write_text("##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\n1\t5\t.\tA\tC\t.\tPASS\t.\tGT\t0/1\n")
    good = "x\nd\nc\nhs37d5\n0\n1\t5\tA\tC\t0.1\t0.2\n"
    for bad in (good.replace("0.1", "0x1p-3"), good.replace("0.1", " 0.1"), good.replace("0.2", "nan(7)"), good.replace("\t5\t", "\t 5\t"),
                good.replace("hs37d5\n0\n", "hs37d5\n 0\n")):
        sc = d / "b.score"
        sc.write_text(bad)
        rc = plan(api, str(sc), str(d / "v.vcf"))[0]
        assert rc == -3, (bad, rc)
    sc = d / "g.score"
    sc.write_text(good.replace("0.2", "nan"))
    assert plan(api, str(sc), str(d / "v.vcf"))[0] == 0


def test_reader_delivers_format_ds(api, tmp_path, monkeypatch):
    """Dosage mode of the readers (FORMAT/DS, one float per sample; not a reference feature): VCF text and BCF
    hand out the same BCF float bits, "." / absent values as the float missing sentinel, and a field placed after
    other FORMAT fields is found."""
    rng = np.random.default_rng(3)
    n = 7
    samples = [f"s{i}" for i in range(n)]
    recs = []
    for k in range(12):
        ds = np.round(rng.uniform(0, 2, n), 3).astype(np.float32)
        ds[rng.random(n) < 0.2] = np.nan
        recs.append(dict(contig="1", pos=100 + 10 * k, ref="A", alts=["C"], filter="PASS", ds=ds,
                         gt=((rng.integers(0, 2, size=(n, 2)) + 1) << 1).astype(np.int8)))
    vcf, bcf = str(tmp_path / "d.vcf.gz"), str(tmp_path / "d.bcf")
    write_vcf(vcf, samples, recs, contigs=["1"], compress="bgzf")
    write_bcf(bcf, samples, recs, ["1"], extra_fmt=True)
    monkeypatch.setenv("NIMPRESS_READ_DS", "1")
    want = np.stack([r["ds"] for r in recs]).copy().view(np.uint32)
    want[~np.isfinite(np.stack([r["ds"] for r in recs]))] = 0x7F800001
    for f in (vcf, bcf):
        got, ns, w, pl = read_gt(api, f, 4 * n)
        assert (ns, w, pl) == (n, 4, 1) and got.shape[0] == len(recs)
        assert np.array_equal(got.view(np.uint32).reshape(len(recs), n), want), f
    monkeypatch.delenv("NIMPRESS_READ_DS")
    got, ns, w, pl = read_gt(api, bcf, 2 * n)                    # GT still there
    assert (w, pl) == (1, 2) and np.array_equal(got.view(np.int8).reshape(len(recs), n, 2), np.stack([r["gt"] for r in recs]))
