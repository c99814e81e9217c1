"""Summarise an .ncu-rep (ncu --set full) into a small JSON: one entry per captured launch with the
metrics the roofline discussion uses.  Usage: python scripts/ncu_summary.py rep.ncu-rep out.json [note]"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {"_note": note, "_source": rep, "launches": []}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"id": d["ID"], "kernel": d["Kernel Name"][:90]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                if r[i] not in ("", "n/a"):
                    e[h] = f"{r[i]} {units[i]}".strip()
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
                try:
                    st[h[len(STALL):-len("_per_issue_active.ratio")]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        e["stall_warps_per_issue_top"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:5])
        res["launches"].append(e)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{len(res['launches'])} launches -> {out}")


if __name__ == "__main__":
    main()
