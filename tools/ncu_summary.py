"""Turn gpurun_out/*.ncu-rep (ncu --set full) and the launch list CSV into the small, committed
summaries under profiles/.  Usage: python tools/ncu_summary.py <round-tag> <rep> [<rep> ...]"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_cmd_atom.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = float(r[i].replace(",", ""))
                except ValueError:
                    d[w] = r[i]
                d[w + ".unit"] = units[i]
        res.append(d)
    return res


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs("profiles", exist_ok=True)
    summary = {}
    lines = ["# ncu --set full summaries (%s)\n" % tag]
    for rep in sys.argv[2:]:
        for idx, k in enumerate(raw(rep)):
            short = k["kernel"].split("(")[0].replace("void ", "").replace("evk::", "")
            name = "%s#%d %s" % (os.path.basename(rep).replace(".ncu-rep", ""), idx, short)
            rd = to_bytes(k.get("dram__bytes_read.sum", 0), k.get("dram__bytes_read.sum.unit", "byte"))
            wr = to_bytes(k.get("dram__bytes_write.sum", 0), k.get("dram__bytes_write.sum.unit", "byte"))
            k["dram_bytes_total"] = rd + wr
            summary[name] = k
            lines.append("## %s\n\n`%s`\n" % (name, k["kernel"][:150]))
            lines.append("| metric | value | unit |\n|---|---|---|")
            for w in WANT:
                if w in k:
                    lines.append("| %s | %s | %s |" % (w, k[w], k.get(w + ".unit", "")))
            lines.append("| dram bytes read+write | %.0f | byte |\n" % k["dram_bytes_total"])
    with open("profiles/ncu_full_%s.md" % tag, "w") as f:
        f.write("\n".join(lines))
    with open("profiles/ncu_full_%s.json" % tag, "w") as f:
        json.dump(summary, f, indent=1)
    print("\n".join(lines))
