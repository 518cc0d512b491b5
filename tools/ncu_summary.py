"""Turn ncu outputs into the small text / JSON summaries committed under profiles/ (runs without a GPU).
    python tools/ncu_summary.py list   launches.csv            -> per-kernel time shares of the launch list
    python tools/ncu_summary.py full   prof.ncu-rep [...]      -> key --set full metrics per profiled launch
"""
import collections, csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]


def short(name):
    name = name.replace("frt2::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
    return name.split("(")[0][:60]


def rows_of(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("\n".join(lines))))


def cmd_list(path):
    rows = rows_of(open(path).read())
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(u, 1.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
        total += v
        n += 1
    print(f"# launches: {n}; sum of kernel durations {total / 1e3:.3f} ms (ncu: cold-cache, serialised — compare SHARES)")
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v / 1e3:10.3f} ms  {100 * v / total:5.1f}%  x{c:4d}  avg {v / c:9.1f} us  {k}")


def cmd_full(paths):
    out = {}
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = rows_of(txt)
        if not rows:
            print("# no rows in", path)
            continue
        for r in rows[1:] if rows[0].get("ID", "") == "" else rows:   # first data row holds the units
            name = short(r.get("Kernel Name", ""))
            if not name:
                continue
            vals = {}
            for k in KEYS:
                if k in r and r[k] not in ("", None):
                    try:
                        vals[k] = float(r[k].replace(",", ""))
                    except ValueError:
                        pass
            units = rows[0]
            dur = vals.get("gpu__time_duration.sum")
            if dur is not None:
                u = units.get("gpu__time_duration.sum", "ns")
                dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
            else:
                dur_us = None

            def to_bytes(key):
                if key not in vals:
                    return None
                return vals[key] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get(key, "byte"), 1)
            rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
            line = (f"{name:34s} grid {int(vals.get('launch__grid_size', 0)):6d} x {int(vals.get('launch__block_size', 0)):4d} "
                    f"regs {int(vals.get('launch__registers_per_thread', 0)):3d} | {dur_us:9.1f} us | "
                    f"DRAM rd {rd / 1e6 if rd is not None else -1:9.1f} MB wr {wr / 1e6 if wr is not None else -1:9.1f} MB "
                    f"({vals.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', -1):5.1f}% of peak) | "
                    f"tensor pipe {vals.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', -1):5.1f}% | "
                    f"issue {vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", -1):5.1f}% | XU {vals.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', -1):5.1f}% | "
                    f"warps {vals.get('sm__warps_active.avg.pct_of_peak_sustained_active', -1):5.1f}% | L2 hit {vals.get('lts__t_sector_hit_rate.pct', -1):5.1f}%")
            print(line)
            out.setdefault(name, []).append({"us": dur_us, "dram_read_bytes": rd, "dram_write_bytes": wr,
                                             "tensor_pipe_pct": vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                             "grid": vals.get("launch__grid_size")})
    return out


if __name__ == "__main__":
    if sys.argv[1] == "list":
        cmd_list(sys.argv[2])
    else:
        res = cmd_full([a for a in sys.argv[2:] if not a.startswith("--json=")])
        js = [a for a in sys.argv[2:] if a.startswith("--json=")]
        if js:
            json.dump(res, open(js[0][7:], "w"), indent=1)
