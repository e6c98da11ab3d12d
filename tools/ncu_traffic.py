"""profiles/r02_traffic.json from an `ncu --set full` capture of ONE update step's tma_gemm launches.

  python tools/ncu_traffic.py RAW_CSV STEP_BREAKDOWN_TXT OUT_JSON

RAW_CSV             `ncu -i <rep> --page raw --csv` of a capture taken with `-k regex:tma_gemm -c 12` (one step)
STEP_BREAKDOWN_TXT  tools/prof_breakdown.py output of the same build: its tma_gemm lines, in launch order, name the layers
bench.py reads the result for roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
import csv
import json
import sys


def main():
    raw, breakdown, out = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    launches = []
    for r in data:
        if "tma_gemm" not in r[col["Kernel Name"]]:
            continue
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        launches.append({"kernel": r[col["Kernel Name"]].split("(")[0], "dram_read": rd, "dram_write": wr,
                         "time": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]})
    labels = []
    for line in open(breakdown):
        parts = line.split()
        if parts and ":tma_gemm" in parts[0] and parts[0] not in labels and not line.startswith("---"):
            labels.append(parts[0])
    # a step launches each label once except the two forwards, which share layer names: phase tells them apart
    per_launch = []
    for i, l in enumerate(launches):
        lab = labels[i] if i < len(labels) else "?"
        per_launch.append(dict(l, label=lab, traffic=l["dram_read"] + l["dram_write"]))
    per_layer = {}
    for l in per_launch:
        if l["label"] == "?":
            continue
        layer = l["label"].split(":")[1]
        per_layer.setdefault(layer, []).append(l["traffic"])
    per_layer = {k: sum(v) / len(v) for k, v in per_layer.items()}
    json.dump({"kernel": launches[0]["kernel"].replace("void ", "").replace("bb::", "") if launches else "",
               "source": raw.split("/")[-1], "n_launches": len(launches), "labels_matched": len(labels) == len(launches),
               "per_layer": per_layer, "per_launch": per_launch}, open(out, "w"), indent=1)
    print("launches", len(launches), "labels", len(labels))
    for l in per_launch:
        print("%-44s rd %9.0f wr %9.0f  %s" % (l["label"], l["dram_read"], l["dram_write"], l["time"]))


if __name__ == "__main__":
    main()
