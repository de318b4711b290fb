"""Summarise ncu reports (gpurun_out/*.ncu-rep, one kernel launch each, `ncu --set full --clock-control none`)
into profiles/*.json:  python tools/ncu_summary.py out.json name=report.ncu-rep [name=report ...]"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read_GB",
    "dram__bytes_write.sum": "dram_write_GB",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_inst",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_ld_sectors",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.max": "sm_cycles",
}
UNIT = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "s": 1e3}


def summarise(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for i, name in enumerate(head):
        if name == "Kernel Name":
            out["kernel"] = vals[i].split("(")[0].replace("void cb::", "")
        if name in KEYS:
            x = float(vals[i].replace(",", ""))
            out[KEYS[name]] = x * UNIT.get(units[i], 1.0)
    if "dram_read_GB" in out:
        out["dram_bytes_per_launch"] = (out["dram_read_GB"] + out["dram_write_GB"]) * 1e9
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        head = rows[1]
        stalls = {}
        for r in rows[2:]:
            for i, k in enumerate(head):
                if k.startswith("stall_") and "Not Issued" not in k and i < len(r) and r[i]:
                    stalls[k] = stalls.get(k, 0) + int(r[i])
        tot = sum(stalls.values()) or 1
        out["warp_stall_samples_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
    return out


if __name__ == "__main__":
    res = {}
    for kv in sys.argv[2:]:
        name, rep = kv.split("=", 1)
        res[name] = summarise(rep)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(res, indent=1))
