"""Tests-only: synthetic score file + VCF/BCF + BED with the corner cases of the reference's
lookup rules (src/nimpress.nim:313-345, 353-364, 375-379, 553)."""
import os

import numpy as np

from util_bcf import write_bcf, write_vcf

BASES = "ACGT"


def make_dataset(tmp, rng, n=300, V=120, miss_rate=0.03, with_traps=True, gt_dtype=np.int8, ploidy=2, haploid_rate=0.0,
                 contigs=("1", "2", "X"), sorted_scores=True, index=False, spread=40):
    """Returns dict(score, vcf, bcf, bed, samples, records, entries).  index: records are put in file order
    (contig, position; same-site records keep their order) and <vcf>.tbi / <bcf>.csi are written.  spread:
    average distance between sites."""
    samples = [f"S{i + 1}" for i in range(n)]
    records = []
    vend = {1: -127, 2: -32767, 4: -2147483647}[np.dtype(gt_dtype).itemsize]
    for c in contigs:
        pos = np.sort(rng.choice(np.arange(100, 100 + spread * V), size=V // len(contigs), replace=False))
        for p in pos:
            ref = BASES[rng.integers(4)] if rng.random() < 0.85 else "".join(BASES[i] for i in rng.integers(0, 4, size=rng.integers(2, 5)))
            n_alt = 1 if rng.random() < 0.8 else int(rng.integers(2, 4))
            alts = []
            while len(alts) < n_alt:
                a = BASES[rng.integers(4)] if rng.random() < 0.85 else "".join(BASES[i] for i in rng.integers(0, 4, size=2))
                if a != ref and a not in alts:
                    alts.append(a)
            af = rng.uniform(0.02, 0.5)
            al = (rng.random((n, ploidy)) < af) * rng.integers(1, n_alt + 1, size=(n, ploidy))
            g = ((al + 1) << 1) | (rng.random((n, ploidy)) < 0.3)
            g[:, 0] &= ~1                                                    # first allele carries no phase bit in text VCF
            miss = rng.random(n) < miss_rate * rng.uniform(0, 3)
            g[miss] = g[miss] & 1
            half = rng.random(n) < 0.004
            g[half, 0] = 0
            if haploid_rate > 0 and ploidy > 1:
                hap = rng.random(n) < haploid_rate
                g[hap, 1:] = vend
            filt = "PASS" if rng.random() < 0.8 else ("." if rng.random() < 0.5 else ("FAIL" if rng.random() < 0.7 else "FAIL;LowQ"))
            records.append(dict(contig=c, pos=int(p), ref=ref, alts=alts, filter=filt, gt=g.astype(gt_dtype)))
    if with_traps:
        c = contigs[0]
        base = max(r["pos"] for r in records if r["contig"] == c) + 100
        g0 = records[0]["gt"]
        # same site twice (split multi-allelic): the FIRST record with REF and ALT matching wins
        records.append(dict(contig=c, pos=base, ref="A", alts=["C"], filter="PASS", gt=g0.copy()))
        records.append(dict(contig=c, pos=base, ref="A", alts=["G"], filter="PASS", gt=np.roll(g0, 1, axis=0)))
        records.append(dict(contig=c, pos=base, ref="A", alts=["G"], filter="FAIL", gt=np.roll(g0, 2, axis=0)))
        # a 2-base REF at base+10 also overlaps a query for the same REF string at base+11 (POS never compared)
        records.append(dict(contig=c, pos=base + 10, ref="AT", alts=["A"], filter="PASS", gt=np.roll(g0, 3, axis=0)))
        # INFO/END extends the record's span: overlaps a query at base+25
        records.append(dict(contig=c, pos=base + 20, ref="G", alts=["T"], filter="PASS", info="END=" + str(base + 30), gt=np.roll(g0, 4, axis=0)))
        # no ALT at all
        records.append(dict(contig=c, pos=base + 40, ref="C", alts=[], filter="PASS", gt=np.full_like(g0, 2)))
    # score entries
    entries = []
    for r in records:
        u = rng.random()
        beta = round(float(rng.normal(0, 0.05)), 4)
        eaf = round(float(rng.uniform(0.01, 0.6)), 4)
        if u < 0.55 and r["alts"]:
            ea = r["alts"][rng.integers(len(r["alts"]))]
            entries.append((r["contig"], r["pos"], r["ref"], ea, beta, eaf))
        elif u < 0.70:
            entries.append((r["contig"], r["pos"], r["ref"], r["ref"], beta, eaf))                 # REF is the effect allele
        elif u < 0.78:
            entries.append((r["contig"], r["pos"], r["ref"], "N", beta, eaf))                      # allele not in ALT: absent
        elif u < 0.84:
            entries.append((r["contig"], r["pos"] + 1, r["ref"], r["ref"], beta, eaf))             # off by one
        elif u < 0.88:
            entries.append(("nochrom", r["pos"], r["ref"], r["ref"], beta, float("nan")))          # unknown contig, NaN eaf
    if with_traps:
        c = contigs[0]
        entries += [(c, base, "A", "G", 0.11, 0.2), (c, base, "A", "C", -0.07, 0.3), (c, base, "A", "T", 0.05, 0.3),
                    (c, base + 11, "AT", "AT", 0.21, 0.4), (c, base + 25, "G", "T", 0.02, 0.1), (c, base + 25, "G", "G", 0.03, 0.9),
                    (c, base + 40, "C", "C", 0.04, 0.99), (c, base + 40, "C", "T", 0.04, 0.01),
                    (c, base, "A", "G", 0.5, float("nan"))]                                         # duplicate entry, NaN eaf
    if not sorted_scores:
        order = rng.permutation(len(entries))
        entries = [entries[i] for i in order]
    score = os.path.join(tmp, "d.score")
    with open(score, "w") as fh:
        fh.write("Synthetic\ndesc\tx  \ncite\nhs37d5\n-0.25\n")
        fh.write("\n".join("\t".join(str(x) for x in e) for e in entries))               # no trailing newline, like wood
    bed = os.path.join(tmp, "d.bed")
    with open(bed, "w") as fh:
        rows = []
        for e in entries:
            u = rng.random()
            stop = e[1] + len(e[2]) - 1
            if e[0] == "nochrom":
                continue
            if u < 0.6:
                rows.append((e[0], e[1] - 1 - int(rng.integers(0, 3)), stop + int(rng.integers(0, 3))))     # covers
            elif u < 0.7:
                rows.append((e[0], e[1], stop + 5))                                                         # start == pos: not covered
            elif u < 0.8:
                rows.append((e[0], e[1] - 5, stop - 1))                                                     # ends one short
        fh.write("\n".join(f"{c}\t{s}\t{t}\textra" for c, s, t in rows))
    if index:
        order = {c: i for i, c in enumerate(contigs)}
        records = sorted(records, key=lambda r: (order[r["contig"]], r["pos"]))      # stable: duplicates keep their order
    vcf = os.path.join(tmp, "d.vcf.gz")
    write_vcf(vcf, samples, records, contigs=list(contigs), filters=("FAIL", "LowQ"), compress="bgzf", index="tbi" if index else None)
    bcf = os.path.join(tmp, "d.bcf")
    write_bcf(bcf, samples, records, contigs=list(contigs), filters=("FAIL", "LowQ"), compress="bgzf", index=bool(index))
    return dict(score=score, vcf=vcf, bcf=bcf, bed=bed, samples=samples, records=records, entries=entries)


def write_score(path, entries, offset=0.0, name="Synthetic"):
    with open(path, "w") as fh:
        fh.write(f"{name}\ndesc\ncite\nhs37d5\n{offset!r}\n")
        fh.write("\n".join("\t".join(str(x) for x in e) for e in entries) + "\n")
    return path


def derive_scores(tmp, rng, entries, S, keep=0.7, flip=0.15):
    """S further score files over (a subset of) the same sites: other weights, some rows with the
    other allele as effect allele, some rows repeated, shuffled order -- what a batch of published
    scores over one cohort looks like."""
    paths = []
    for k in range(S):
        ents = []
        for (c, pos, ref, ea, beta, eaf) in entries:
            if rng.random() > keep:
                continue
            if rng.random() < flip:
                ea = ref if ea != ref else "N"
            ents.append((c, pos, ref, ea, round(float(rng.normal(0, 0.05)), 4), eaf))
            if rng.random() < 0.02:
                ents.append(ents[-1])
        if k % 2:
            ents = [ents[i] for i in rng.permutation(len(ents))]
        paths.append(write_score(os.path.join(tmp, f"derived{k}.score"), ents, offset=round(float(rng.normal()), 3), name=f"derived{k}"))
    return paths
