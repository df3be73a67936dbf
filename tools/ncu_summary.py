"""Summarise ncu CSV exports (raw + source pages) brought back from the GPU box.  usage: ncu_summary.py <prefix>"""
import collections
import csv
import sys

pre = sys.argv[1]
rows = list(csv.reader(open(f"{pre}_raw.csv")))
r = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def _bytes(key):
    v, u = r[key]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


if "--traffic-json" in sys.argv:      # usage: ncu_summary.py <prefix> --traffic-json <out.json> <side> <source text>
    import json
    i = sys.argv.index("--traffic-json")
    side = int(sys.argv[i + 2])
    json.dump({"kernel": r.get("Kernel Name", ("?",))[0], "side": side,
               "dram_bytes_read": _bytes("dram__bytes_read.sum"), "dram_bytes_write": _bytes("dram__bytes_write.sum"),
               "gpu_time": r["gpu__time_duration.sum"][0] + " " + r["gpu__time_duration.sum"][1],
               "warp_instructions": float(r["smsp__inst_executed.sum"][0]),
               "thread_instr_per_pixel": float(r["smsp__inst_executed.sum"][0]) * 32.0 / (side * side),
               "issue_active_pct": float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "tensor_pipe_active_pct": float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
               "registers_per_thread": float(r["launch__registers_per_thread"][0]),
               "source": sys.argv[i + 3]}, open(sys.argv[i + 1], "w"), indent=1)
print("kernel:", r.get("Kernel Name", ("?",))[0])
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum ",
        "sm__pipe_tensor_subpipe", "smsp__warp_issue_stalled"]
for k in sorted(r):
    if any(k.startswith(x) for x in keys) and not k.endswith(".per_second") and "pct_of_peak_sustained_elapsed" not in k[40:]:
        print(f"  {k:84s} {r[k][0]} {r[k][1]}")
rows = list(csv.reader(open(f"{pre}_source.csv")))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, samp, stalls, tot, totsamp = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
lines = []
for rr in rows[2:]:
    try:
        c, s = int(rr[ia]), int(rr[isamp])
    except Exception:
        continue
    txt = rr[isrc].strip()
    parts = txt.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    ops[op] += c; samp[op] += s; tot += c; totsamp += s
    for i in stall_cols:
        try:
            stalls[hdr[i]] += int(rr[i])
        except Exception:
            pass
    lines.append((s, c, txt))
print(f"total warp-instr {tot}  samples {totsamp}  SASS lines {len(lines)}")
for op, c in ops.most_common(24):
    print(f"  {op:12s} {c:12d} {100 * c / tot:5.1f}%   samples {100 * samp[op] / max(1, totsamp):5.1f}%")
print("  stalls:", [(k, round(100 * v / max(1, totsamp), 1)) for k, v in stalls.most_common(9)])
print("hottest SASS lines by samples:")
for s, c, txt in sorted(lines, reverse=True)[:14]:
    print(f"  {100 * s / max(1, totsamp):5.1f}%  exec {c:10d}  {txt[:100]}")
