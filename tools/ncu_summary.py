"""Summarise an .ncu-rep of one kernel: headline metrics, stall mix, dynamic opcode mix and the execution-count
profile of the SASS (which loop a region belongs to).  usage: python tools/ncu_summary.py report.ncu-rep [--sass lo hi]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
for k in keys:
    if k in m: print(f"{k:85s} {m[k]:>14s} {u[k]}")
print("stalls per issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_')[0]} {float(v):.2f}" for k, v in sorted(m.items(), key=lambda kv: -float(kv[1] or 0) if 'smsp__average_warps_issue_stalled' in kv[0] and 'not_issued' not in kv[0] else 0) if 'smsp__average_warps_issue_stalled' in k and 'not_issued' not in k and float(v or 0) > 0.05))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try: data.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[ix["Source"]].strip()))
    except Exception: pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("warp instructions", tot, "samples", ts)
op = collections.Counter(); ops = collections.Counter()
for e, s, t in data:
    k = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0]
    op[k] += e; ops[k] += s
print("opcode mix:", ", ".join(f"{k} {100*v/tot:.1f}%/{100*ops[k]/ts:.1f}%s" for k, v in op.most_common(16)))
unit = data[0][0]
prev = None; start = 0
print(f"execution profile (x = executions per warp; first instruction ran {unit} times):")
for i, (e, s, t) in enumerate(data + [(-1, 0, "")]):
    k = round(e / unit, 2)
    if k != prev:
        if prev is not None and prev > 0.3:
            print(f"  [{start}-{i-1}] x{prev} n={i-start} instr/warp={prev*(i-start):.0f} samples={100*sum(d[1] for d in data[start:i])/ts:.1f}%")
        prev = k; start = i
if "--sass" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--sass") + 1]), int(sys.argv[sys.argv.index("--sass") + 2])
    for i in range(lo, hi + 1): print(i, data[i][1], data[i][2])
