"""Tests-only host logic in Python: read a small VCF, apply the reference's coverage / lookup /
FILTER rules (src/nimpress.nim:313-345, 353-364, 553) and emit one block in the C-ABI's input
form (raw int8 GT rows + npc_row array).  The product's host is the C++ one; this exists so the
golden vectors can be driven through the CUDA path independently of it."""
import gzip

import numpy as np

import orc


def read_vcf(path):
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    samples, recs = [], []
    with op(path, "rt", newline="") as fh:
        for ln in fh.read().replace("\r\n", "\n").split("\n"):
            if not ln or ln.startswith("##"):
                continue
            f = ln.split("\t")
            if ln.startswith("#"):
                samples = f[9:]
                continue
            fmt = f[8].split(":")
            gi = fmt.index("GT")
            gts = []
            for s in f[9:9 + len(samples)]:
                sub = s.split(":")
                g = sub[gi] if gi < len(sub) else "."
                al, phased = [], 0
                tok = ""
                for ch in g + "\0":
                    if ch in "/|\0":
                        al.append(phased if tok in (".", "") else ((int(tok) + 1) << 1) | phased)
                        phased = 1 if ch == "|" else 0
                        tok = ""
                    else:
                        tok += ch
                gts.append(al)
            recs.append(dict(contig=f[0], pos=int(f[1]), ref=f[3], alt=[] if f[4] == "." else f[4].split(","),
                             filter=f[6], gts=gts))
    return samples, recs


def read_score(path):
    lines = open(path).read().split("\n")
    offset = float(lines[4].rstrip())
    ents = []
    for ln in lines[5:]:
        if ln == "":
            continue
        c, p, r, e, b, a = ln.rstrip().split("\t")
        ents.append(dict(contig=c, pos=int(p), ref=r, ea=e, beta=float(b), eaf=float(a)))
    return offset, ents


def read_bed(path):
    iv = {}
    for ln in open(path).read().split("\n"):
        if ln == "":
            continue
        c, s, e = ln.rstrip().split("\t")[:3]
        iv.setdefault(c, []).append((int(s), int(e)))
    return iv


def build_block(score_path, vcf_path, bed_path=None, ignorefilt=False):
    samples, recs = read_vcf(vcf_path)
    offset, ents = read_score(score_path)
    bed = read_bed(bed_path) if bed_path else None
    n = len(samples)
    stride = -(-2 * n // 16) * 16
    gt = np.zeros((len(recs), stride), dtype=np.int8)
    for i, r in enumerate(recs):
        for s, al in enumerate(r["gts"]):
            al = (al + [-127, -127])[:2]            # vector_end padding for short calls
            gt[i, 2 * s], gt[i, 2 * s + 1] = al
    rows = np.zeros(len(ents), dtype=orc.ROW_DTYPE)
    for j, e in enumerate(ents):
        stop = e["pos"] + len(e["ref"]) - 1
        rows[j]["beta"], rows[j]["eaf"] = e["beta"], e["eaf"]
        rows[j]["ref_is_ea"] = int(e["ref"] == e["ea"])
        rows[j]["gt_row"], rows[j]["eaidx"] = -1, -1
        if bed is not None and not any(s < e["pos"] and t >= stop for s, t in bed.get(e["contig"], [])):
            rows[j]["kind"] = 1
            continue
        hit = None
        for i, r in enumerate(recs):
            if r["contig"] != e["contig"] or r["pos"] > stop or r["pos"] + len(r["ref"]) - 1 < e["pos"]:
                continue
            if r["ref"] == e["ref"] and (e["ea"] == e["ref"] or e["ea"] in r["alt"]):
                hit = i
                break
        if hit is None:
            rows[j]["kind"] = 2
            continue
        r = recs[hit]
        rows[j]["gt_row"] = hit
        rows[j]["eaidx"] = 0 if e["ea"] == r["ref"] else r["alt"].index(e["ea"]) + 1
        rows[j]["kind"] = 3 if (not ignorefilt and r["filter"] not in (".", "PASS")) else 0
    return samples, offset, gt, rows
