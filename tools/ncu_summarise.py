"""Summarise an `ncu --set full` report: one row per profiled launch with the numbers the
rooflines are checked against.

    python tools/ncu_summarise.py gpurun_out/x.ncu-rep [--labels probe.json] [--out profiles/x.json]

--labels: the JSON list tools/hbm_probe.py prints (launch order == profile order, matched by
position among the launches whose kernel name passes --match): adds algorithmic bytes and the
traffic / algorithmic ratio.
"""
import argparse
import csv
import io
import json
import subprocess

COLS = {
    "duration_ns": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct_of_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct_of_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l2_to_sm_bytes": "lts__t_bytes_srcunit_tex.sum",
    "sm_warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3,
              "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--labels")
    ap.add_argument("--match", default="", help="only kernels whose name contains this")
    ap.add_argument("--skip", default="", help="comma list of substrings: kernels to drop (e.g. torch fills)")
    ap.add_argument("--out")
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    skip = [s for s in args.skip.split(",") if s]
    out = []
    for r in data:
        name = r[idx["Kernel Name"]]
        if args.match and args.match not in name:
            continue
        if any(s in name for s in skip):
            continue
        e = {"kernel": name, "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k, col in COLS.items():
            if col not in idx:
                continue
            v = num(r[idx[col]])
            if v is None:
                continue
            v *= UNIT_SCALE.get(units[idx[col]], 1.0)
            e[k] = v
        if "duration_ns" in e:
            e["duration_us"] = e.pop("duration_ns") / 1e3
            tr = e.get("dram_read_bytes", 0.0) + e.get("dram_write_bytes", 0.0)
            e["dram_gbs"] = tr / (e["duration_us"] * 1e-6) / 1e9
        out.append(e)
    if args.labels:
        labels = json.load(open(args.labels))
        if len(labels) != len(out):
            raise SystemExit("labels (%d) and profiled launches (%d) differ; use --match / --skip" % (len(labels), len(out)))
        for e, lb in zip(out, labels):
            e["launch"] = lb["launch"]
            e["algorithmic_bytes"] = lb["algorithmic_bytes"]
            tr = e.get("dram_read_bytes", 0.0) + e.get("dram_write_bytes", 0.0)
            e["traffic_over_algorithmic"] = tr / lb["algorithmic_bytes"]
            e["algorithmic_gbs"] = lb["algorithmic_bytes"] / (e["duration_us"] * 1e-6) / 1e9
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    for e in out:
        print("%-64s %8.1f us  dram %7.1f+%7.1f MB  %6.0f GB/s  dram%% %5.1f  tensor%% %5.1f  %s" % (
            e.get("launch", e["kernel"].split("(")[0][-64:]), e.get("duration_us", 0), e.get("dram_read_bytes", 0) / 1e6,
            e.get("dram_write_bytes", 0) / 1e6, e.get("dram_gbs", 0), e.get("dram_throughput_pct", 0),
            e.get("tensor_pipe_pct_of_active", 0),
            ("x%.2f of algorithmic" % e["traffic_over_algorithmic"]) if "traffic_over_algorithmic" in e else ""))


if __name__ == "__main__":
    main()
