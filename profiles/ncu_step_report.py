"""Per-kernel table of an `ncu --set full` capture of one step replay (profiles/prof_step.py), read without a GPU:
    python profiles/ncu_step_report.py gpurun_out/r02_step_yelp.ncu-rep > profiles/r02_ncu_full_step_C2.txt
Columns: duration, DRAM bytes read / written, L2 (lts) bytes, grid, registers, warps active, issue slots busy, FMA pipe,
LSU pipe, shared-memory wavefronts (% of peak) and the three largest warp-stall reasons (pc sampling)."""
import csv
import json
import re
import subprocess
import sys


def main(path, traffic_json=None, workload=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale=1.0, default=float("nan")):
        if name not in col:
            return default
        try:
            return float(r[col[name]].replace(",", "")) * scale
        except ValueError:
            return default

    def unit(name):
        return units[col[name]] if name in col else ""

    def to_bytes(r, name):
        u = unit(name).lower()
        mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return get(r, name, mul, 0.0)

    stall_cols = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    print(f"{'kernel':38} {'us':>7} {'dram rd MB':>10} {'dram wr MB':>10} {'L2 MB':>8} {'grid':>6} {'regs':>5} {'warps%':>7} "
          f"{'issue%':>7} {'fma%':>6} {'lsu%':>6} {'smem wf%':>8}  top stall reasons (pc samples)")
    out = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")[:38]
        dur = get(r, "gpu__time_duration.sum", {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit("gpu__time_duration.sum"), 1.0))
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        l2 = to_bytes(r, "lts__t_bytes.sum")
        stalls = sorted(((get(r, c, 1.0, 0.0), c.replace("smsp__pcsamp_warps_issue_stalled_", "")) for c in stall_cols), reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        top = ", ".join(f"{n} {v / tot * 100:.0f}%" for v, n in stalls[:3])
        print(f"{name:38} {dur:7.2f} {rd / 1e6:10.3f} {wr / 1e6:10.3f} {l2 / 1e6:8.2f} {get(r, 'launch__grid_size'):6.0f} "
              f"{get(r, 'launch__registers_per_thread'):5.0f} {get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{get(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{get(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{get(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):8.1f}  {top}")
        out[name] = {"us": dur, "dram_bytes": rd + wr}
    if traffic_json:
        grp = lambda names: sum(v["dram_bytes"] for k, v in out.items() if any(k.startswith(n) for n in names))
        try:
            t = json.load(open(traffic_json))
        except Exception:
            t = {}
        t[workload] = {"choose": {"bytes": grp(["k_choose_wide", "k_choose_small", "k_choose_huge", "k_choose_big"]),
                                  "kernels": "k_choose_wide + k_choose_small (+ huge / big tiers)"},
                       "choose_prep": {"bytes": grp(["k_choose_prep"]), "kernels": "k_choose_prep"},
                       "aggregate": {"bytes": grp(["k_aggregate"]), "kernels": "k_aggregate"},
                       "source": path.split("/")[-1] + " (ncu --set full, one step replay, L2 flushed before it)"}
        json.dump(t, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
