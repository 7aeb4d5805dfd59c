"""`ncu --set full` capture of the training step's kernels -> profiles/r02_ncu_summary.json.

    ncu -i gpurun_out/step_r2.ncu-rep --page raw --csv > /tmp/step_raw.csv
    python tools/ncu_step_summary.py /tmp/step_raw.csv profiles/r02_bench_final.json > profiles/r02_ncu_summary.json

ncu does not know the plan's labels, so each captured launch is matched to the bench line's `step_kernels`
entries with the same kernel family AND grid (two layers with identical shapes share one entry: their
launches are averaged).  Rows: {"kernel": "<label(s)> [family]", "grid", "launches", "time_us",
"dram_read_bytes", "dram_write_bytes", "l2_to_sm_bytes", "lts_pct", "dram_pct", "tensor_pipe_pct",
"warp_inst"} -- per launch.  bench.py's `roofline.traffic` = dram_read_bytes + dram_write_bytes of the row
whose "kernel" contains the roofline entry's label.
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from step_roofline import family, read_bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    raw, bench = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in col or r[col[name]] in ("", "n/a", "no data"):
            return None
        try:
            return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)
        except ValueError:
            return None

    labels = {}
    for k in read_bench(bench)["step_kernels"]:
        key = (family(k["kernel"]), "x".join(str(g) for g in k.get("grid", [])))
        labels.setdefault(key, []).append((k["ms"], k["label"]))
    METRICS = (("time_us", "gpu__time_duration.sum"), ("dram_read_bytes", "dram__bytes_read.sum"),
               ("dram_write_bytes", "dram__bytes_write.sum"), ("l2_to_sm_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
               ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
               ("warp_inst", "smsp__inst_executed.sum"))
    launches = {}
    for r in rows[2:]:
        grid = r[col["Grid Size"]].strip("() ").replace(" ", "").replace(",", "x")
        key = (family(r[col["Kernel Name"]]), grid)
        if key in labels:
            launches.setdefault(key, []).append({name: (val(r, metric) or 0.0) for name, metric in METRICS})
    out = []
    for key, ls in launches.items():
        names = sorted(labels[key], reverse=True)                 # longest first (bench event timing)
        ls.sort(key=lambda a: -a["time_us"])                      # longest first (ncu)
        # several layers can share one kernel family + grid (the persistent weight-gradient kernel always
        # launches one CTA per SM): pair them up by duration rank when the counts allow it
        if len(names) > 1 and len(ls) % len(names) == 0:
            per = len(ls) // len(names)
            groups = [([names[i][1]], ls[i * per:(i + 1) * per]) for i in range(len(names))]
        else:
            groups = [(sorted(set(n for _, n in names)), ls)]
        for lab, g in groups:
            row = {"kernel": "%s [%s]" % (" / ".join(lab), key[0]), "grid": key[1], "launches": len(g)}
            for name, _ in METRICS:
                row[name] = sum(a[name] for a in g) / len(g)
            out.append(row)
    out.sort(key=lambda r: -r["time_us"] * r["launches"])
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
