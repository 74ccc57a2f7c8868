"""profiles/rNN_march_ncu_summary.json from an .ncu-rep of the march kernel -- usage: ncu_summary.py file.ncu-rep key "capture command" [algorithmic bytes] [round, default r02]"""
import csv, json, os, subprocess, sys
rep, key, how = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
def num(name, scale_by_unit=True):
    v = float(m[name].replace(",", ""))
    if scale_by_unit:
        v *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3}.get(u[name], 1)
    return v
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
s = {
    "capture": how, "kernel": m.get("Kernel Name", "k_fused_march"),
    "gpu_time_us": num("gpu__time_duration.sum"), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "algorithmic_bytes_per_launch": int(sys.argv[4]) if len(sys.argv) > 4 else 132710400,
    "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
    "warp_instructions": num("smsp__inst_executed.sum", False),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
    "pipe_fma_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
    "pipe_alu_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", False),
    "pipe_xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False),
    "pipe_lsu_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", False),
    "registers_per_thread": num("launch__registers_per_thread", False), "grid": num("launch__grid_size", False), "block": num("launch__block_size", False),
    "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
    "shared_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False),
    "l2_hit_pct": num("lts__t_sector_hit_rate.pct", False),
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", (sys.argv[5] if len(sys.argv) > 5 else "r02") + "_march_ncu_summary.json")
j = json.load(open(path)) if os.path.exists(path) else {}
j[key] = s
json.dump(j, open(path, "w"), indent=1)
print(json.dumps(s, indent=1))
