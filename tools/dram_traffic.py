"""ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the launches of one decode
step) -> the per-kernel DRAM traffic JSON bench.py reads for `roofline.traffic` (profiles/rNN_dram_traffic.json).
usage: python tools/dram_traffic.py launches_dram.csv "source description" > profiles/r02_dram_traffic.json"""
import collections, json, sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from ncu_summary import rows_of, short  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}


def main():
    rows = rows_of(open(sys.argv[1]).read())
    per = collections.OrderedDict()
    for r in rows:
        k = (r["ID"], short(r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        per.setdefault(k, {})[r["Metric Name"]] = v
    agg = collections.OrderedDict()
    for (_id, name), m in per.items():
        a = agg.setdefault(name, {"launches": 0, "dram_read_gb": 0.0, "dram_write_gb": 0.0, "us": 0.0})
        a["launches"] += 1
        a["dram_read_gb"] += m.get("dram__bytes_read.sum", 0.0) / 1e9
        a["dram_write_gb"] += m.get("dram__bytes_write.sum", 0.0) / 1e9
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
    for a in agg.values():
        a["traffic_gb_per_launch"] = (a["dram_read_gb"] + a["dram_write_gb"]) / a["launches"]
        a["avg_gbs"] = (a["dram_read_gb"] + a["dram_write_gb"]) / (a["us"] * 1e-6) if a["us"] else None
    print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "kernels": agg}, indent=1))


if __name__ == "__main__":
    main()
