#!/usr/bin/env python
"""Digest of an .ncu-rep (raw page + source page) as one small JSON: key counters per profiled launch, the warp-stall mix, DMMA issued
against DMMA predicated-on (a predicated-off DMMA still occupies the FP64 pipe), and the hottest SASS windows.  Run where the report
is (the GPU box: reports with --import-source are tens of MB and do not travel):  ncu_digest.py report.ncu-rep out.json"""
import collections
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__shared_mem_per_block_dynamic"]


def page(rep, name):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout.splitlines()))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = page(rep, "raw")
    h = raw[0]
    digest = {"report": rep, "launches": []}
    for r in raw[2:]:
        d = {w: r[h.index(w)] for w in WANT if w in h}
        d["kernel"] = r[h.index("Kernel Name")][:120] if "Kernel Name" in h else ""
        d["units"] = {w: raw[1][h.index(w)] for w in WANT if w in h}
        digest["launches"].append(d)
    src = page(rep, "source")
    # the source page is one table per launch: "Kernel Name" line, header line, rows
    tables, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"kernel": r[1][:120], "hdr": None, "rows": []}
            tables.append(cur)
        elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
            cur["rows"].append(r)
    digest["source"] = []
    for t in tables:
        hd, rows = t["hdr"], t["rows"]
        ix = {k: i for i, k in enumerate(hd)}
        num = lambda r, k: float(r[ix[k]] or 0)
        tot = sum(num(r, "# Samples") for r in rows) or 1.0
        st = collections.Counter()
        for r in rows:
            for k in hd:
                if k.startswith("stall_") and "Not Issued" not in k:
                    st[k] += num(r, k)

        def op(r):
            s = r[ix["Source"]].split()
            if not s:
                return ""
            return (s[1] if s[0].startswith("@") else s[0]).split(".")[0]

        dmma_issued = sum(num(r, "Instructions Executed") for r in rows if op(r) == "DMMA")
        dmma_on = sum(num(r, "Predicated-On Thread Instructions Executed") for r in rows if op(r) == "DMMA") / 32.0
        ops = collections.Counter()
        for r in rows:
            ops[op(r)] += num(r, "# Samples")
        w, hot = 48, []
        for a in range(0, len(rows), w):
            seg = rows[a:a + w]
            s = sum(num(r, "# Samples") for r in seg)
            hot.append((s, a, seg))
        hot.sort(key=lambda x: -x[0])
        hot_out = []
        for s, a, seg in hot[:12]:
            oc = collections.Counter(op(r) for r in seg)
            hot_out.append({"sass_index": a, "samples_pct": round(100 * s / tot, 2),
                            "stalls_pct": {k[6:]: round(100 * sum(num(r, k) for r in seg) / tot, 2) for k in ("stall_long_sb", "stall_math", "stall_wait", "stall_short_sb", "stall_barrier")},
                            "ops": dict(oc.most_common(5))})
        digest["source"].append({"kernel": t["kernel"], "sass_instructions": len(rows), "samples": tot,
                                 "stall_mix_pct": {k[6:]: round(100 * v / tot, 1) for k, v in st.most_common(10)},
                                 "samples_by_opcode_pct": {k: round(100 * v / tot, 1) for k, v in ops.most_common(10)},
                                 "dmma_warp_instructions_issued": dmma_issued, "dmma_warp_instructions_predicated_on": dmma_on,
                                 "hot_windows": hot_out})
    json.dump(digest, open(out, "w"), indent=1)
    print("digest", out, len(digest["launches"]), "launches")


if __name__ == "__main__":
    main()
