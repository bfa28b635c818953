#!/usr/bin/env python3
"""BASELINE.json configs[3] from DISK through the command line: 18 score files (the bundled wood-height
definition + 17 derived over the same ~700 sites: other weights, some with REF as effect allele) on ONE
synthetic 200,000-sample BCF (BGZF).  `nimpress a,b,...,r genotypes.bcf` (one pass over the file, one
resident slab, tensor-core contraction) against 18 runs of `nimpress <one file> genotypes.bcf` -- which
is how the reference would be used.  Every file's printed scores are checked against the oracle.

    python tools/bench_cli_config4.py [--samples 200000] [--out gpurun_out/cli_config4.json]
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
from util_files import write_score           # noqa: E402
from util_vcf import read_score              # noqa: E402

SEED = 0x6E696D70


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200_000)
    ap.add_argument("--scores", type=int, default=18)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_config4.json"))
    args = ap.parse_args()
    n, S = args.samples, args.scores
    wood = os.path.join(ROOT, "tests", "golden", "scores", "wood-25282103-height.scores")
    offset, ents = read_score(wood)
    rng = np.random.default_rng(SEED)
    sites = [dict(contig=e["contig"], pos=e["pos"], ref=e["ref"], alts=[e["ea"] if e["ea"] != e["ref"] else "N"], af=e["eaf"]) for e in ents]
    contigs = sorted({s["contig"] for s in sites}, key=lambda c: (len(c), c))
    order = sorted(range(len(sites)), key=lambda i: (contigs.index(sites[i]["contig"]), sites[i]["pos"]))
    row_of = {j: i for i, j in enumerate(order)}                    # score entry -> record index in the file
    V = len(sites)
    gt = np.zeros((V, 2 * n), np.int8)
    af = np.array([min(max(sites[j]["af"], 0.01), 0.99) for j in order])
    orc.synth_fill(gt, n, 0, SEED, (af * 65536).astype(np.uint32), np.full(V, int(0.005 * (1 << 24)), np.uint32), np.ones(V, np.int32))
    samples = [f"S{i:06d}" for i in range(n)]
    tmp = tempfile.mkdtemp(prefix="npcli4_")
    bcf = os.path.join(tmp, "config4.bcf")
    write_bcf(bcf, samples, [dict(contig=sites[j]["contig"], pos=sites[j]["pos"], ref=sites[j]["ref"], alts=sites[j]["alts"], filter="PASS",
                                  gt=gt[i].reshape(n, 2)) for i, j in enumerate(order)], contigs, filters=())
    # the definitions: file 0 = wood itself
    paths, rows_list, offs = [wood], [], [offset]
    base = np.zeros(len(ents), dtype=orc.ROW_DTYPE)
    for j, e in enumerate(ents):
        base[j] = (row_of[j], 0 if e["ea"] == e["ref"] else 1, e["beta"], e["eaf"], int(e["ea"] == e["ref"]), 0)
    rows_list.append(base)
    for k in range(1, S):
        keep = rng.random(len(ents)) < 0.9
        es, rows = [], []
        for j, e in enumerate(ents):
            if not keep[j]:
                continue
            beta = round(float(rng.normal(0, 0.05)), 4)
            ea = e["ea"]
            if rng.random() < 0.2 and e["ea"] != e["ref"]:
                ea = e["ref"]                                         # this definition counts the REF allele
            es.append((e["contig"], e["pos"], e["ref"], ea, beta, e["eaf"]))
            rows.append((row_of[j], 0 if ea == e["ref"] else 1, beta, e["eaf"], int(ea == e["ref"]), 0))
        off = round(float(rng.normal()), 3)
        paths.append(write_score(os.path.join(tmp, f"derived{k}.score"), es, offset=off, name=f"derived{k}"))
        rows_list.append(np.array(rows, dtype=orc.ROW_DTYPE)); offs.append(off)
    want = [orc.score_matrix(gt, n, 2, rows_list[k], offset=offs[k], threads=os.cpu_count() or 1)["scores"] for k in range(S)]

    exe = os.path.join(ROOT, "nimpress_b200", "bin", "nimpress")

    def parse(stdout):
        blocks, cur = [], None
        for l in stdout.splitlines():
            if l.startswith("#score"):
                cur = []; blocks.append(cur)
            elif not l.startswith("WARN"):
                if cur is None:
                    cur = []; blocks.append(cur)
                cur.append(float(l.split("\t")[1]))
        return [np.array(b) for b in blocks]

    def check(got, k):
        w = want[k]
        assert len(got) == n, (k, len(got))
        dev = np.abs(got - w) / np.maximum(np.abs(w), 1e-3)
        assert np.all(dev <= 1e-12), f"file {k} differs from the oracle: max dev {np.nanmax(dev):.3e} at {int(np.nanargmax(dev))}, nan {int(np.isnan(got).sum())}/{int(np.isnan(w).sum())}, got {got[:3]} want {w[:3]}"

    res = {"workload": f"config4: {S} score files x {V} loci x {n} samples, BCF/BGZF on disk", "bcf_bytes": os.path.getsize(bcf),
           "gt_bytes": int(V) * 2 * n, "genotype_score_cells": int(sum(len(r) for r in rows_list)) * n, "host_cores": os.cpu_count()}
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        p = subprocess.run([exe, "--afmisp=0", ",".join(paths), bcf], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr
        best = dt if best is None else min(best, dt)
    for k, b in enumerate(parse(p.stdout)):
        check(b, k)
    res["one_pass_all_files_wall_s"] = best
    t0 = time.perf_counter()
    for k in range(S):
        p = subprocess.run([exe, "--afmisp=0", paths[k], bcf], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        check(parse(p.stdout)[0], k)
    res["one_run_per_file_wall_s"] = time.perf_counter() - t0
    res["ratio"] = res["one_run_per_file_wall_s"] / res["one_pass_all_files_wall_s"]
    env = dict(os.environ, NIMPRESS_TIMING="1", NPC_TIMING="1")
    p = subprocess.run([exe, "--afmisp=0", ",".join(paths), bcf], capture_output=True, text=True, env=env)
    res["phases"] = [l for l in p.stderr.splitlines() if l.startswith("[")]
    res["scores_match_oracle"] = True
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
