#!/usr/bin/env python3
"""Regenerate tests/golden/ from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_fixtures.py

What it writes (all DATA, no reference source code):

* ``set1/``   - byte copies of the reference's own test inputs
                (``tests/set1.{score,bed,vcf.gz,vcf.gz.tbi}``, ``tests/set1.plink190.result``),
                needed because the parity tests must run where the reference tree is absent.
* ``scores/`` - byte copies of the four bundled score DEFINITIONS (``scores/*.scores``); the
                other 14 files in ``scores/`` are result files.
* ``res_format_corpus.tsv`` - the 3,528 ``sample<TAB>score`` lines of the 14 bundled result
                files with the sample names replaced by an index: a corpus of the reference's
                float output format (``%.16g`` + ``.0`` rule).
* ``set1_expected.json`` - the 13 expected score vectors of ``tests/test_set1.nim`` and
                ``stats_kat.json`` the known-answer values of ``tests/test_stats.nim``,
                transcribed by regex from those files, each with its file:line.
"""
import json, os, re, shutil, sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def copy(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")
    for f in ("set1.score", "set1.bed", "set1.vcf.gz", "set1.vcf.gz.tbi",
              "set1.plink190.result", "set1.plink190.result.txt"):
        copy(f"{REF}/tests/{f}", f"{HERE}/set1/{f}")
    for f in sorted(os.listdir(f"{REF}/scores")):
        if f.endswith(".scores"):
            copy(f"{REF}/scores/{f}", f"{HERE}/scores/{f}")

    # output-format corpus
    lines = []
    for f in sorted(os.listdir(f"{REF}/scores")):
        if f.endswith("_nimpress_res.txt"):
            for ln in open(f"{REF}/scores/{f}"):
                ln = ln.rstrip("\n")
                if ln:
                    lines.append(ln.split("\t")[1])
    with open(f"{HERE}/res_format_corpus.tsv", "w") as fh:
        fh.write("\n".join(lines) + "\n")

    # test_set1.nim expectations
    src = open(f"{REF}/tests/test_set1.nim").read().split("\n")
    cases, i = [], 0
    enum = lambda s: s.split(".")[-1].strip().rstrip(",")
    while i < len(src):
        m = re.match(r'\s*test "(.*)":', src[i])
        if m and not src[i].lstrip().startswith("#"):
            name, line0 = m.group(1), i + 1
            body = []
            i += 1
            while i < len(src) and "check(checkFloats" not in src[i]:
                body.append(src[i]); i += 1
            chk_line = i + 1
            body_s = " ".join(body)
            call = re.search(r"computePolygenicScores\(scores, scoreFile, genotypeVcf, (\w+), coveredBed,(.*)\)", body_s)
            cov = call.group(1) == "true"
            rest = call.group(2)
            loc = enum(re.search(r"ImputeMethodLocus\.(\w+)", rest).group(0))
            mis = enum(re.search(r"ImputeMethodMissing\.(\w+)", rest).group(0))
            smp = enum(re.search(r"ImputeMethodSample\.(\w+)", rest).group(0))
            maxmis = float(re.search(r"maxMissingRate\s*=\s*([0-9.]+)", rest).group(1))
            afp = float(re.search(r"afMismatchPthresh\s*=\s*([0-9.]+)", rest).group(1))
            mincs = int(re.search(r"minGtForInternalImput\s*=\s*([0-9]+)", rest).group(1))
            ign = re.search(r"ignoreFilterField\s*=\s*(\w+)", rest).group(1) == "true"
            vec = re.search(r"@\[(.*)\]", src[i]).group(1)
            exp = []
            for tok in vec.split(","):
                tok = tok.strip()
                if tok == "NaN":
                    exp.append(None)
                elif "-" in tok[1:] and not tok.startswith("-"):   # "0.123-0.03"
                    a, b = tok.split("-")
                    exp.append(["sub", float(a), float(b)])
                else:
                    exp.append(float(tok))
            cases.append(dict(name=name, ref=f"tests/test_set1.nim:{line0}-{chk_line}", cov=cov,
                              imp_locus=loc, imp_missing=mis, imp_sample=smp, maxmis=maxmis,
                              afmisp=afp, mincs=mincs, ignorefilt=ign, expected=exp,
                              tolerance_abs=1e-4))
        i += 1
    assert len(cases) == 13, len(cases)
    json.dump(dict(source="tests/test_set1.nim (checkFloats tolerance 1e-4 abs, NaN pattern exact; :14-22)",
                   cases=cases), open(f"{HERE}/set1_expected.json", "w"), indent=1)

    # test_stats.nim KATs
    kats = []
    for ln_no, ln in enumerate(open(f"{REF}/tests/test_stats.nim"), 1):
        m = re.search(r"check_floatvalue\((\w+)\(([^)]*)\),\s*([0-9.eE+-]+)\)", ln)
        if m:
            kats.append(dict(fn=m.group(1), args=[float(a) for a in m.group(2).split(",")],
                             target=float(m.group(3)), kind="approx", ref=f"tests/test_stats.nim:{ln_no}"))
            continue
        m = re.search(r"^\s*(\w+)\(([^)]*)\)\s*==\s*([0-9.]+)\s*$", ln)
        if m:
            kats.append(dict(fn=m.group(1), args=[float(a) for a in m.group(2).split(",")],
                             target=float(m.group(3)), kind="exact", ref=f"tests/test_stats.nim:{ln_no}"))
    json.dump(dict(source="tests/test_stats.nim (rel 1e-5, abs 1e-9 when |target|<1e-9; :6-17)", kats=kats),
              open(f"{HERE}/stats_kat.json", "w"), indent=1)
    print(f"{len(cases)} set1 cases, {len(kats)} stats KATs, {len(lines)} format-corpus floats")


if __name__ == "__main__":
    main()
