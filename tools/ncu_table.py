"""One line per profiled launch of an `ncu --set full` report (run here, no GPU needed): duration, DRAM bytes, DRAM / L2 / tensor
utilisation, issue-slot use -- the evidence table for profiles/ and DESIGN.md.

    python tools/ncu_table.py gpurun_out/prof.ncu-rep [--json key] > profiles/r02_kernels_x.txt
"""
import csv
import io
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rdMB", 1e-6), ("dram__bytes_write.sum", "wrMB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smemLSU%", 1),
        ("launch__grid_size", "grid", 1), ("launch__registers_per_thread", "regs", 1)]
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, as_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    print(f"{'kernel':34s} " + " ".join(f"{n:>9s}" for _, n, _ in COLS))
    agg = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:34]
        vals = []
        for key, _, scale in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                vals.append(float("nan")); continue
            v = float(r[ix[key]].replace(",", "")) * UNIT_SCALE.get(units[ix[key]], 1.0)
            vals.append(v * scale)
        print(f"{name:34s} " + " ".join(f"{v:9.2f}" for v in vals))
        agg.setdefault(name, []).append(vals)
    if as_json:
        d = {k: {"launches": len(v), "us": sum(x[0] for x in v) / len(v), "dram_bytes": int(sum((x[1] + x[2]) for x in v) / len(v) * 1e6)}
             for k, v in agg.items()}
        print(json.dumps({as_json: d}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
