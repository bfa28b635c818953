#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/:
key raw metrics per kernel launch, the warp-stall breakdown, and -- from the source page (needs
-lineinfo / --import-source) -- the executed-instruction mix and shared-memory wavefronts.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--base N] > profiles/<name>.txt

--base N: divide instruction counts by N (e.g. rows x CTAs x consumer warps) to get per-unit costs."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    base = float(sys.argv[sys.argv.index("--base") + 1]) if "--base" in sys.argv else None
    rows = ncu(rep, "raw")
    h, u = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        d = dict(zip(h, r))
        units = dict(zip(h, u))
        print(f"\n## {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k]:>18s} {units[k]}")
        st = [(k[33:], float(v.replace(",", ""))) for k, v in d.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and v]
        tot = sum(v for _, v in st) or 1
        print("warp stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st, key=lambda x: -x[1])[:8]))
    src = ncu(rep, "source")
    if len(src) > 2 and "Source" in src[1]:
        hh, data = src[1], src[2:]
        isrc, iex = hh.index("Source"), hh.index("Instructions Executed")
        iw = hh.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hh else None
        iwi = hh.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hh else None
        mix, wf, wfi = collections.Counter(), collections.Counter(), collections.Counter()
        for r in data:
            if len(r) <= iex:
                continue
            s = r[isrc].strip()
            if not s:
                continue
            op = s.split()[1] if s.startswith("@") and len(s.split()) > 1 else s.split()[0]
            if not (r[iex] or "0").isdigit():
                continue                                  # repeated header of the next kernel's listing
            ex = int(r[iex] or 0)
            mix[op.split(".")[0]] += ex
            if iw is not None and r[iw]:
                wf[op] += int(r[iw] or 0); wfi[op] += int(r[iwi] or 0)
        tot = sum(mix.values())
        div = base or 1.0
        unit = "per unit (--base)" if base else "total"
        print(f"\n## executed warp-instructions (all kernels of the report), {unit}: {tot / div:.1f}")
        print(", ".join(f"{k} {v / div:.2f}" for k, v in mix.most_common(24)))
        if wf:
            print(f"## shared-memory wavefronts {unit} (actual / ideal): total {sum(wf.values()) / div:.2f} / {sum(wfi.values()) / div:.2f}")
            print(", ".join(f"{k} {v / div:.2f}/{wfi[k] / div:.2f}" for k, v in wf.most_common(8)))


if __name__ == "__main__":
    main()
