#!/usr/bin/env python3
"""BASELINE.json configs[1] from DISK through the command line: the 697 wood-height loci on a
synthetic 100,000-sample BCF (BGZF), `nimpress <scoredef> <genotypes.bcf>` end to end -- file read,
BGZF inflate, BCF record walk, streaming findVariant, staging, H2D, kernels, output formatting.
Checks the printed scores against the oracle (exact-order mode: every printed digit) and reports
wall time with the sequential inflater and with the BGZF worker pool.

    python tools/bench_cli_config2.py [--samples 100000] [--out gpurun_out/cli_config2.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc                                   # noqa: E402  (checker only)
from util_bcf import write_bcf               # noqa: E402
from util_vcf import read_score              # noqa: E402

SEED = 0x6E696D70


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_config2.json"))
    ap.add_argument("--timing", action="store_true", help="also print the CLI's phase timing (NIMPRESS_TIMING=1) for one run of each mode")
    args = ap.parse_args()
    n = args.samples
    score = os.path.join(ROOT, "tests", "golden", "scores", "wood-25282103-height.scores")
    offset, ents = read_score(score)
    rng = np.random.default_rng(SEED)
    # the 697 score sites + 10% decoy sites, sorted by contig / position like a real BCF
    sites = [dict(contig=e["contig"], pos=e["pos"], ref=e["ref"], alts=[e["ea"] if e["ea"] != e["ref"] else "N"], score_idx=i, af=e["eaf"])
             for i, e in enumerate(ents)]
    for _ in range(len(ents) // 10):
        e = ents[rng.integers(len(ents))]
        sites.append(dict(contig=e["contig"], pos=e["pos"] + int(rng.integers(50, 5000)), ref="A", alts=["G"], score_idx=-1, af=0.2))
    contigs = sorted({s["contig"] for s in sites}, key=lambda c: (len(c), c))
    sites.sort(key=lambda s: (contigs.index(s["contig"]), s["pos"]))
    V = len(sites)
    stride = 2 * n
    gt = np.zeros((V, stride), np.int8)
    af = np.array([min(max(s["af"], 0.01), 0.99) if s["af"] == s["af"] else 0.3 for s in sites])
    orc.synth_fill(gt, n, 0, SEED, (af * 65536).astype(np.uint32), np.full(V, int(0.005 * (1 << 24)), np.uint32), np.ones(V, np.int32))
    # ref == ea rows count REF alleles: nothing to change in the genotypes
    samples = [f"S{i:06d}" for i in range(n)]
    tmp = tempfile.mkdtemp(prefix="npcli_")
    bcf = os.path.join(tmp, "config2.bcf")
    t0 = time.perf_counter()
    write_bcf(bcf, samples, [dict(contig=s["contig"], pos=s["pos"], ref=s["ref"], alts=s["alts"], filter="PASS",
                                  gt=gt[i].reshape(n, 2)) for i, s in enumerate(sites)], contigs, filters=())
    t_write = time.perf_counter() - t0
    # oracle on the same genotypes (in-memory entry): rows in score-file order
    rows = np.zeros(len(ents), dtype=orc.ROW_DTYPE)
    row_of = {s["score_idx"]: i for i, s in enumerate(sites) if s["score_idx"] >= 0}
    for j, e in enumerate(ents):
        rows[j] = (row_of[j], 0 if e["ea"] == e["ref"] else 1, e["beta"], e["eaf"], int(e["ea"] == e["ref"]), 0)
    want = orc.score_matrix(gt, n, 2, rows, offset=offset)
    want_txt = [orc.format_float(x) for x in want["scores"]]

    exe = os.path.join(ROOT, "nimpress_b200", "bin", "nimpress")
    res = {"workload": f"config2: {len(ents)} wood loci (+{V - len(ents)} decoy records) x {n} samples, BCF/BGZF on disk",
           "bcf_bytes": os.path.getsize(bcf), "gt_bytes": int(V) * stride, "genotypes": len(ents) * n, "runs": []}
    for threads, extra in (("1", ["--exact-order"]), ("", ["--exact-order"]), ("", [])):
        env = dict(os.environ)
        if threads:
            env["NIMPRESS_THREADS"] = threads
        else:
            env.pop("NIMPRESS_THREADS", None)
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "--afmisp=0", *extra, score, bcf], capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            assert p.returncode == 0, p.stderr
            best = dt if best is None else min(best, dt)
        lines = [l for l in p.stdout.splitlines() if not l.startswith("WARN")]
        got_txt = [l.split("\t")[1] for l in lines]
        assert [l.split("\t")[0] for l in lines] == samples
        if "--exact-order" in extra:
            assert got_txt == want_txt, "printed scores differ from the oracle"
        else:
            g, w = np.array([float(x) for x in got_txt]), want["scores"]
            assert np.all(np.abs(g - w) <= 1e-12 * np.maximum(np.abs(w), 1e-3))
        res["runs"].append({"inflate_threads": threads or "default (min(cores,16))", "mode": "exact-order" if extra else "tile4 (default)",
                            "wall_s_best_of_3": best, "genotypes_per_s_from_disk": len(ents) * n / best,
                            "scores_match_oracle": True})
    if args.timing:
        for extra in (["--exact-order"], []):
            for th in ("1", "16"):
                env = dict(os.environ, NIMPRESS_TIMING="1", NIMPRESS_THREADS=th)
                p = subprocess.run([exe, "--afmisp=0", *extra, score, bcf], capture_output=True, text=True, env=env)
                print(f"--- {extra} threads={th}\n" + p.stderr, file=sys.stderr)
    res["bcf_write_s_python"] = t_write
    res["host_cores"] = os.cpu_count()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
