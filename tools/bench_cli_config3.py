#!/usr/bin/env python3
"""A slice of BASELINE.json configs[2] from DISK through the command line ("end-to-end via BCF is
host/PCIe-bound and reported separately on a 10k-variant slice"): V variants x 500,000 samples of the synthetic
genome-wide cohort written as one BGZF BCF, then `nimpress <scoredef> <genotypes.bcf>` end to end -- file read,
BGZF inflate (the engine's own DEFLATE decoder on a thread pool), BCF record walk, streaming findVariant,
pinned staging, H2D into the resident slab, the fused kernel, output formatting.  The printed scores are
checked against the oracle on the same genotypes.

    python tools/bench_cli_config3.py [--variants 10000] [--samples 500000] [--devices 0-1]
"""
import argparse
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc                                   # noqa: E402  (checker only)
from bench import SEED, cohort_params        # noqa: E402
from util_bcf import _desc, _typed_int, _typed_ints, _typed_str      # noqa: E402

BLOCK = 0xFF00


def bgzf_block(chunk):
    co = zlib.compressobj(1, zlib.DEFLATED, -15)
    comp = co.compress(chunk) + co.flush()
    return (struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(comp) + 25) + comp +
            struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def write_bcf_streaming(path, n, V, contig, positions, gt_rows_iter, pool):
    """BCF2.2 / BGZF written in pieces: header, then records (REF A, ALT C, PASS, GT int8 diploid) whose GT rows come from
    gt_rows_iter in batches; the BGZF blocks of a batch are deflated on the pool (zlib releases the GIL)."""
    samples = [f"S{i:06d}" for i in range(n)]
    lines = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed">', f"##contig=<ID={contig}>",
             '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
             "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + samples)]
    text = ("\n".join(lines) + "\n").encode() + b"\0"
    pending = bytearray(b"BCF\2\2" + struct.pack("<I", len(text)) + text)
    gt_key = _typed_int(1)                                         # dictionary: PASS = 0, GT = 1
    total = 0
    with open(path, "wb") as f:
        def flush(final=False):
            nonlocal pending, total
            nfull = len(pending) // BLOCK * BLOCK if not final else len(pending)
            view = bytes(pending[:nfull])
            chunks = [view[i:i + BLOCK] for i in range(0, nfull, BLOCK)]
            for b in pool.map(bgzf_block, chunks):
                f.write(b); total += len(b)
            pending = pending[nfull:]
        k = 0
        for rows in gt_rows_iter:
            for r in range(rows.shape[0]):
                shared = struct.pack("<iiif", 0, int(positions[k]) - 1, 1, float("nan")) + struct.pack("<II", (2 << 16) | 0, (1 << 24) | n)
                shared += _desc(0, 7) + _typed_str("A") + _typed_str("C") + _typed_ints([0])
                indiv_hdr = gt_key + _desc(2, 1)
                pending += struct.pack("<II", len(shared), len(indiv_hdr) + 2 * n) + shared + indiv_hdr
                pending += rows[r, :2 * n].tobytes()
                k += 1
            flush()
        flush(final=True)
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    assert k == V
    return total + 28


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=10_000)
    ap.add_argument("--samples", type=int, default=500_000)
    ap.add_argument("--devices", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_config3.json"))
    args = ap.parse_args()
    n, V = args.samples, args.variants
    af, beta, ref_is_ea, af_thr, miss_thr, alt = cohort_params(0, V)
    positions = 10_000 + 137 * np.arange(V)
    tmp = tempfile.mkdtemp(prefix="npcli3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    bcf, score = os.path.join(tmp, "slice.bcf"), os.path.join(tmp, "slice.score")
    with open(score, "w") as f:
        f.write("config3 slice\nsynthetic\ncitation\nhs37d5\n0.125\n")
        for v in range(V):
            f.write(f"1\t{positions[v]}\tA\t{'A' if ref_is_ea[v] else 'C'}\t{float(beta[v])!r}\t{round(float(af[v]), 4)!r}\n")
    threads = os.cpu_count() or 4
    pool = ThreadPoolExecutor(threads)
    stride = 2 * n
    batch = max(1, (256 << 20) // stride)
    # oracle expectation accumulated batch by batch (raw sums are not exposed: keep the batches' normalised scores * 2 * nloci)
    rows_all = np.zeros(V, dtype=orc.ROW_DTYPE)
    rows_all["eaidx"] = np.where(ref_is_ea == 1, 0, 1); rows_all["beta"] = beta; rows_all["eaf"] = np.round(af, 4)
    rows_all["ref_is_ea"] = ref_is_ea
    want_sum = np.zeros(n)
    want_nloci = 0

    def batches():
        nonlocal want_sum, want_nloci
        for v0 in range(0, V, batch):
            nb_ = min(batch, V - v0)
            gt = np.zeros((nb_, stride), np.int8)
            orc.synth_fill(gt, n, v0, SEED, af_thr[v0:v0 + nb_], miss_thr[v0:v0 + nb_], alt[v0:v0 + nb_])
            r = rows_all[v0:v0 + nb_].copy()
            r["gt_row"] = np.arange(nb_)
            o = orc.score_matrix(gt, n, 2, r, offset=0.0, threads=threads)
            want_sum += o["scores"] * (2.0 * o["nloci"])
            want_nloci += o["nloci"]
            yield gt
    t0 = time.perf_counter()
    size = write_bcf_streaming(bcf, n, V, "1", positions, batches(), pool)
    t_write = time.perf_counter() - t0
    want = want_sum / (2.0 * want_nloci) + 0.125
    floor = orc.abs_floor(beta, want_nloci)

    exe = os.path.join(ROOT, "nimpress_b200", "bin", "nimpress")
    res = {"workload": f"config3 slice: {V} variants x {n} samples, one BGZF BCF on {'tmpfs' if tmp.startswith('/dev/shm') else 'disk'}",
           "bcf_bytes": size, "gt_bytes": int(V) * stride, "genotypes": V * n, "host_cores": threads, "runs": []}
    extra = [f"--devices={args.devices}"] if args.devices else []
    for rep in range(3):
        env = dict(os.environ, NIMPRESS_TIMING="1")
        t0 = time.perf_counter()
        p = subprocess.run([exe, "--afmisp=0", *extra, score, bcf], capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-2000:]
        got = np.array([float(l.split("\t")[1]) for l in p.stdout.splitlines() if not l.startswith("WARN")])
        assert got.shape == (n,)
        # the oracle total above is itself a batch-wise re-association: compare at the re-association tolerance, twice over
        err = np.abs(got - want) / (1e-12 * np.abs(want) + 2 * floor)
        assert err.max() <= 2.0, err.max()
        phases = {ln.split("]")[1].rsplit(None, 2)[0].strip(): float(ln.rsplit(None, 2)[1]) for ln in p.stderr.splitlines() if ln.startswith("[nimpress timing]") and ln.rstrip().endswith("ms")}
        res["runs"].append({"wall_s": dt, "genotypes_per_s_from_disk": V * n / dt, "gt_payload_gb_per_s": V * stride / dt / 1e9,
                            "phases_ms": phases, "scores_match_oracle": True, "devices": args.devices or "0"})
    res["bcf_write_s_python"] = t_write
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))
    for f in (bcf, score):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
