"""profiles/ncu_traffic.json from `ncu --set full` captures: measured DRAM bytes per launch of every kernel.
   python scripts/ncu_traffic.py [--fresh] <label>=<raw.csv>[@suffix] [...]   (raw.csv = `ncu -i X.ncu-rep --page raw --csv`)
Later files override earlier ones per kernel; a kernel already taken from an earlier file of the same call is stored as
<kernel>@<suffix> instead (the shared front-end / NES kernels of the C3 capture next to the C2 one).  The label is recorded
as the capture the numbers came from; --fresh starts from an empty table."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    args = [a for a in sys.argv[1:] if a != "--fresh"]
    out = json.load(open(out_path)) if os.path.exists(out_path) and "--fresh" not in sys.argv else {}
    seen = set()
    for arg in args:
        label, path = arg.split("=", 1)
        suffix = ""
        if "@" in os.path.basename(path):
            path, suffix = path.rsplit("@", 1)
        rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}

        def val(r, name):
            v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "%": 1}.get(u, 1)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            if name in seen and suffix:
                name = name + "@" + suffix
            seen.add(name)
            out[name] = {"capture": label, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                         "ncu_duration_us": round(val(r, "gpu__time_duration.sum"), 3),
                         "tensor_pipe_active_pct": round(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 2),
                         "fp64_pipe_active_pct": round(val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 2),
                         "dram_bytes_per_launch": int(rd + wr)}
    json.dump(out, open(out_path, "w"), indent=1)
    print(out_path, len(out), "kernels")


if __name__ == "__main__":
    main()
