"""Tests-only writers: VCF text and BCF2.2 (plain or BGZF) from one in-memory record list, so the
product's readers can be checked against each other and against the oracle's VCF-text reader."""
import struct
import zlib

import numpy as np

VEND = {1: -127, 2: -32767, 4: -2147483647}
INDEX_BLOCK = 0xFF00          # uncompressed bytes per BGZF block of the indexed files; tests shrink it to get many blocks


def bgzf_compress(data, block=0xFF00, offsets=None):
    """offsets: optional list that receives the file offset of every block (for the index writers)."""
    out = bytearray()
    for i in range(0, max(len(data), 1), block):
        if offsets is not None:
            offsets.append(len(out))
        chunk = data[i:i + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = len(comp) + 25
        out += struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize) + comp
        out += struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    out += bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")      # EOF block
    return bytes(out)


def gt_text(row, width):
    """raw BCF GT values of one sample -> VCF GT string"""
    s = ""
    for k, v in enumerate(row):
        v = int(v)
        if v == VEND[width]:
            break
        if k:
            s += "|" if (v & 1) else "/"
        a = (v >> 1) - 1
        s += "." if a < 0 else str(a)
    return s or "."


def write_vcf(path, samples, records, contigs=None, filters=("FAIL",), crlf=False, compress=None, index=None):
    """records: dicts contig,pos,ref,alts(list),filter(str),gt(int array [n,ploidy], BCF encoding), info(str)."""
    nl = "\r\n" if crlf else "\n"
    lines = ["##fileformat=VCFv4.2"]
    lines += [f'##FILTER=<ID={f},Description="x">' for f in filters]
    lines += [f"##contig=<ID={c}>" for c in (contigs or [])]
    lines += ['##INFO=<ID=END,Number=1,Type=Integer,Description="End">',
              '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
              '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">',
              '##FORMAT=<ID=DS,Number=A,Type=Float,Description="Expected ALT dosage">']
    lines.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    for r in records:
        g = np.asarray(r["gt"])
        width = g.dtype.itemsize
        fmt = r.get("format", "GT")
        if "ds" in r:                                      # GT:DS, floats printed so that float32(text) is the value ("." = missing)
            cols = [r["contig"], str(r["pos"]), ".", r["ref"], ",".join(r["alts"]) if r["alts"] else ".", ".", r["filter"], r.get("info", "."), "GT:DS"]
            cols += [gt_text(g[i], width) + ":" + ("." if not np.isfinite(r["ds"][i]) else repr(float(np.float32(r["ds"][i])))) for i in range(len(samples))]
            lines.append("\t".join(cols))
            continue
        cols = [r["contig"], str(r["pos"]), ".", r["ref"], ",".join(r["alts"]) if r["alts"] else ".", ".", r["filter"],
                r.get("info", "."), fmt]
        if fmt == "GT":
            cols += [gt_text(g[i], width) for i in range(len(samples))]
        else:                                              # GT:DP
            cols += [gt_text(g[i], width) + ":7" for i in range(len(samples))]
        lines.append("\t".join(cols))
    data = (nl.join(lines) + nl).encode()
    if index:                                              # tabix index of the BGZF file: needs the byte span of every record line
        assert compress == "bgzf"
        head = len((nl.join(lines[:len(lines) - len(records)]) + nl).encode())
        spans, pos = [], head
        for ln in lines[len(lines) - len(records):]:
            n = len(ln.encode()) + len(nl)
            spans.append((pos, pos + n)); pos += n
        offs = []
        comp = bgzf_compress(data, block=INDEX_BLOCK, offsets=offs)
        open(path, "wb").write(comp)
        write_index(path + (".csi" if index == "csi" else ".tbi"), records, spans, offs, contigs or sorted({r["contig"] for r in records}),
                    csi=index == "csi", tabix_meta=True, block=INDEX_BLOCK)
        return
    if compress == "bgzf":
        data = bgzf_compress(data)
    elif compress == "gzip":
        import gzip
        data = gzip.compress(data)
    open(path, "wb").write(data)


def _typed_int(v):
    if -120 <= v <= 127:
        return struct.pack("<Bb", 0x11, v)
    if -32000 <= v <= 32767:
        return struct.pack("<Bh", 0x12, v)
    return struct.pack("<Bi", 0x13, v)


def _desc(n, t):
    if n < 15:
        return bytes([(n << 4) | t])
    return bytes([0xF0 | t]) + _typed_int(n)


def _typed_str(s):
    b = s.encode()
    return _desc(len(b), 7) + b


def _typed_ints(vals):
    if not len(vals):
        return bytes([0x00])
    lo, hi = min(vals), max(vals)
    if -120 <= lo and hi <= 127:
        return _desc(len(vals), 1) + struct.pack(f"<{len(vals)}b", *vals)
    if -32000 <= lo and hi <= 32767:
        return _desc(len(vals), 2) + struct.pack(f"<{len(vals)}h", *vals)
    return _desc(len(vals), 3) + struct.pack(f"<{len(vals)}i", *vals)


def write_bcf(path, samples, records, contigs, filters=("FAIL",), compress="bgzf", with_idx=False, extra_fmt=False, index=None):
    """BCF2.2.  Dictionary of strings: PASS, then FILTERs, INFO END, FORMAT GT, FORMAT DP (header order,
    or explicit IDX= when with_idx, deliberately permuted)."""
    ids = ["PASS"] + list(filters) + ["END", "GT", "DP", "DS"]
    idx = {s: i for i, s in enumerate(ids)}
    if with_idx:                                           # shuffle the dictionary through IDX=
        perm = [0] + list(range(len(ids) - 1, 0, -1))
        idx = {s: perm[i] for i, s in enumerate(ids)}
    cidx = {c: i for i, c in enumerate(contigs)}
    tag = (lambda s: f",IDX={idx[s]}") if with_idx else (lambda s: "")
    lines = ["##fileformat=VCFv4.2", f'##FILTER=<ID=PASS,Description="All filters passed"{tag("PASS")}>']
    lines += [f'##FILTER=<ID={f},Description="x"{tag(f)}>' for f in filters]
    lines += [f"##contig=<ID={c}>" for c in contigs]
    lines += [f'##INFO=<ID=END,Number=1,Type=Integer,Description="End"{tag("END")}>',
              f'##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype"{tag("GT")}>',
              f'##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth"{tag("DP")}>',
              f'##FORMAT=<ID=DS,Number=A,Type=Float,Description="Expected ALT dosage"{tag("DS")}>']
    lines.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    text = ("\n".join(lines) + "\n").encode() + b"\0"
    out = bytearray(b"BCF\2\2" + struct.pack("<I", len(text)) + text)
    n = len(samples)
    spans = []
    for r in records:
        rec_start = len(out)
        g = np.ascontiguousarray(r["gt"])
        alleles = [r["ref"]] + list(r["alts"])
        rlen = len(r["ref"])
        info = b""
        n_info = 0
        if r.get("info", ".").startswith("END="):
            end = int(r["info"][4:])
            rlen = end - r["pos"] + 1
            info = _typed_int(idx["END"]) + _typed_ints([end])
            n_info = 1
        flt = [] if r["filter"] == "." else [idx[f] for f in r["filter"].split(";")]
        shared = struct.pack("<iiif", cidx[r["contig"]], r["pos"] - 1, rlen, float("nan"))
        n_fmt = (2 if extra_fmt else 1) + (1 if "ds" in r else 0)
        shared += struct.pack("<II", (len(alleles) << 16) | n_info, (n_fmt << 24) | n)
        shared += _desc(0, 7)                              # ID: empty string
        for a in alleles:
            shared += _typed_str(a)
        shared += _typed_ints(flt) + info
        indiv = b""
        if extra_fmt:                                      # a field before GT: the reader must skip it by size
            indiv += _typed_int(idx["DP"]) + _desc(1, 1) + bytes([7] * n)
        t = {1: 1, 2: 2, 4: 3}[g.dtype.itemsize]
        indiv += _typed_int(idx["GT"]) + _desc(g.shape[1], t) + g.tobytes()
        if "ds" in r:                                      # one float per sample; NaN in the input = BCF float missing
            d = np.asarray(r["ds"], dtype=np.float32).copy()
            bits = d.view(np.uint32)
            bits[~np.isfinite(d)] = 0x7F800001
            indiv += _typed_int(idx["DS"]) + _desc(1, 5) + bits.tobytes()
        out += struct.pack("<II", len(shared), len(indiv)) + shared + indiv
        spans.append((rec_start, len(out)))
    data = bytes(out)
    if index:                                              # CSI index (what `bcftools index` writes for a BCF)
        assert compress == "bgzf"
        offs = []
        comp = bgzf_compress(data, block=INDEX_BLOCK, offsets=offs)
        open(path, "wb").write(comp)
        write_index(path + ".csi", records, spans, offs, contigs, csi=True, tabix_meta=False, block=INDEX_BLOCK)
        return
    if compress == "bgzf":
        data = bgzf_compress(data)
    open(path, "wb").write(data)


# ---- tabix / CSI index writers (tests only; the binning scheme of the SAM / tabix / CSI specifications) ----

def _rec_span0(r):
    """0-based half-open reference span of a record: REF length, or INFO/END."""
    beg = r["pos"] - 1
    end = beg + len(r["ref"])
    if r.get("info", ".").startswith("END="):
        end = max(int(r["info"][4:]), r["pos"])
    return beg, end


def _reg2bin(beg, end, min_shift, depth):
    end -= 1
    s, t = min_shift, ((1 << (depth * 3)) - 1) // 7
    for l in range(depth, 0, -1):
        if beg >> s == end >> s:
            return t + (beg >> s)
        s += 3
        t -= 1 << ((l - 1) * 3)
    return 0


def write_index(path, records, spans, block_offsets, contigs, csi, tabix_meta, block=0xFF00, min_shift=14, depth=5):
    """records must be sorted by (contig order, pos) as in the data file.  spans: uncompressed byte span of each record."""
    voff = lambda p: (block_offsets[p // block] << 16) | (p % block) if p // block < len(block_offsets) else ((block_offsets[-1] + 1) << 16)
    per_ref = {c: dict(bins={}, lin={}) for c in contigs}
    for r, (a, b) in zip(records, spans):
        beg, end = _rec_span0(r)
        R = per_ref[r["contig"]]
        bn = _reg2bin(beg, end, min_shift, depth)
        ch = R["bins"].setdefault(bn, [])
        if ch and ch[-1][1] == voff(a):
            ch[-1] = (ch[-1][0], voff(b))                   # adjacent records of one bin: one chunk
        else:
            ch.append((voff(a), voff(b)))
        for w in range(beg >> min_shift, ((end - 1) >> min_shift) + 1):
            R["lin"].setdefault(w, voff(a))                 # first record overlapping the window
    out = bytearray()
    names = b"".join(c.encode() + b"\0" for c in contigs)
    meta = struct.pack("<6i", 2, 1, 2, 0, ord("#"), 0) + struct.pack("<i", len(names)) + names
    if csi:
        aux = meta if tabix_meta else b""
        out += b"CSI\1" + struct.pack("<iii", min_shift, depth, len(aux)) + aux + struct.pack("<i", len(contigs))
    else:
        out += b"TBI\1" + struct.pack("<i", len(contigs)) + meta
    for c in contigs:
        R = per_ref[c]
        nwin = max(R["lin"]) + 1 if R["lin"] else 0
        lin, prev = [], 0
        for w in range(nwin):                               # empty windows carry the previous offset (htslib back-fill)
            prev = R["lin"].get(w, prev)
            lin.append(prev)
        out += struct.pack("<i", len(R["bins"]))
        for bn in sorted(R["bins"]):
            out += struct.pack("<I", bn)
            if csi:                                         # loffset: linear offset of the bin's first window
                lvl, t = 0, 0
                while bn >= t + (1 << (3 * lvl)):
                    t += 1 << (3 * lvl); lvl += 1
                w0 = (bn - t) << (3 * (depth - lvl))
                out += struct.pack("<Q", lin[min(w0, nwin - 1)] if nwin else 0)
            out += struct.pack("<i", len(R["bins"][bn]))
            for a, b in R["bins"][bn]:
                out += struct.pack("<QQ", a, b)
        if not csi:
            out += struct.pack("<i", nwin) + b"".join(struct.pack("<Q", v) for v in lin)
    open(path, "wb").write(bgzf_compress(bytes(out)))
