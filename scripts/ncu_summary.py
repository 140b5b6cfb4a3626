"""Summary JSON of one kernel launch in an .ncu-rep (ncu --set full): the counters profiles/*_ncu.json quote.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "kernel label" "how it was captured" [algorithmic_bytes] > profiles/x_ncu.json
"""
import csv, io, json, subprocess, sys

rep, label, source = sys.argv[1], sys.argv[2], sys.argv[3]
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(head, units, vals)}


def f(name, scale=1.0):
    v, u = m.get(name, ("", ""))
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    mult = {"Kbyte/block": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
            "ns": 1e-3, "nsecond": 1e-3}.get(u, 1.0)
    return x * mult * scale


out = {
    "kernel": label, "source": source,
    "gpu_time_us": f("gpu__time_duration.sum"),
    "sm_clock_ghz": f("sm__cycles_elapsed.avg.per_second"),
    "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
    "dram_cycles_active_pct": f("dram__cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "algorithmic_bytes": alg,
    "warp_instructions": f("smsp__inst_executed.sum") or f("sm__inst_executed.sum"),
    "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "pipe_fma_cycles_active_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_alu_cycles_active_pct": f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_tensor_cycles_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "l1tex_data_pipe_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_ld_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "l2_to_sm_bytes": f("l1tex__m_xbar2l1tex_read_bytes.sum"),
    "registers_per_thread": f("launch__registers_per_thread"),
    "grid": f("launch__grid_size"), "block": f("launch__block_size"),
    "dynamic_smem_bytes": f("launch__shared_mem_per_block_dynamic"),
}
stalls = {}
for h in head:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = f(h)
        if v and v > 0.1:
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
if alg and out["dram_bytes_read"] is not None:
    out["dram_traffic_over_algorithmic"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / alg
print(json.dumps(out, indent=1))
