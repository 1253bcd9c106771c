"""Condenses `ncu -i X.ncu-rep --page raw --csv` (+ optionally `--page source --csv`) of ONE kernel launch into a small
JSON record for profiles/: duration, DRAM bytes, instructions, issue-active, pipe utilisation, stall mix, and -- with
the source page -- the opcode mix.  Usage: python tools/ncu_summary.py raw.csv [source.csv] > summary.json"""
import collections, csv, json, sys

def raw_record(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    body = [r for r in rows[1:] if len(r) == len(hdr) and r[0] not in ("", "ID") and not r[0].startswith("=")]
    vals = body[-1] if body else rows[-1]
    units = rows[1] if len(rows) > 2 else [""] * len(hdr)
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    def f(k):
        try: return float(d[k])
        except Exception: return None
    def scaled(k):                      # bytes / time come with a unit column
        v = f(k)
        if v is None: return None
        mul = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u.get(k, ""), 1.0)
        return v * mul
    rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"),
           "gpu_time_ms_under_ncu": scaled("gpu__time_duration.sum"),
           "dram_bytes_read": scaled("dram__bytes_read.sum"), "dram_bytes_write": scaled("dram__bytes_write.sum"),
           "inst_executed": f("smsp__inst_executed.sum"),
           "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_per_scheduler": f("smsp__warps_active.avg.per_cycle_active"),
           "warps_eligible_per_scheduler": f("smsp__warps_eligible.avg.per_cycle_active"),
           "registers_per_thread": f("launch__registers_per_thread"),
           "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
           "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
           "sm_throughput_pct": f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
           "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
           "shared_atomics": f("smsp__inst_executed_op_shared_atom.sum"),
           "pipes_pct_of_peak_active": {k.split("pipe_")[1].split(".")[0]: f(k) for k in hdr
                                        if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") and (f(k) or 0) >= 0.5},
           "tensor_pipe_pct": {k.split("sm__")[1].split(".")[0]: f(k) for k in hdr
                               if ("pipe_tensor" in k or "pipe_tc" in k or "tmem" in k or "utcmma" in k.lower()) and k.endswith("pct_of_peak_sustained_active") and (f(k) or 0) > 0},
           "stall_per_issue": {k.split("stalled_")[1].replace("_per_issue_active.ratio", ""): round(f(k), 2) for k in hdr
                               if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and (f(k) or 0) >= 0.1}}
    return rec

def opcode_mix(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    data = rows[rows.index(hdr) + 1:]
    iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, smp = collections.Counter(), collections.Counter()
    for r in data:
        try: n, s = int(r[iI]), int(r[iSm])
        except Exception: continue
        m = r[iS].strip().split()
        if not m: continue
        op = (m[0] if not m[0].startswith("@") else m[1]).split(".")[0]
        ops[op] += n; smp[op] += s
    tot, ts = sum(ops.values()) or 1, sum(smp.values()) or 1
    return {op: {"inst_share_pct": round(100.0 * n / tot, 1), "stall_sample_share_pct": round(100.0 * smp[op] / ts, 1)} for op, n in ops.most_common(14)}

rec = raw_record(sys.argv[1])
if len(sys.argv) > 2:
    try: rec["opcode_mix"] = opcode_mix(sys.argv[2])
    except Exception as exc: rec["opcode_mix"] = {"error": repr(exc)}
print(json.dumps(rec, indent=1))
