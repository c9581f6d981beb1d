"""Writes profiles/traffic.json -- measured DRAM traffic per launch of the kernels bench.py reports a roofline for -- from
the raw ncu exports committed beside it (ncu -i X.ncu-rep --page raw --csv > profiles/<name>_raw.csv).

    python profiles/make_traffic.py

bench.py looks a kernel up by (kernel, workload); a kernel / workload without a committed capture gets traffic = null."""
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (raw export, kernel name as bench.py spells it, workload key)
CAPTURES = [
    ("r02_ncu_persist_c4_raw.csv", "k_fused_persist<3,1,0>", "c4 n=119", "k_fused_persist"),
    ("r02_ncu_fused_plain_c4_raw.csv", "k_fused_assemble<3,1,1,0>", "c4 n=119", "k_fused_assemble"),
    ("r02_ncu_cg_raw.csv", "k_spmv_sell<1,1>", "c4 n=119", "k_spmv_sell"),
    ("r02_ncu_fused_c2_raw.csv", "k_fused_assemble<2,1,1,0>", "c2 N=1414", "k_fused_assemble"),
    ("r02_ncu_persist_c3_raw.csv", "k_fused_persist<2,2,0,1>", "c3 N=1000", "k_fused_persist"),
    ("r02_ncu_fused_c3_raw.csv", "k_fused_assemble<2,2,0,1>", "c3 N=1000", "k_fused_assemble"),
    ("r02_ncu_fused_p2tet_raw.csv", "k_fused_assemble<3,2,1,3>", "p2tet n=76", "k_fused_assemble"),
]


def fnum(x):
    return float(x.replace(",", ""))


def main():
    out = []
    for fname, kernel, workload, match in CAPTURES:
        path = os.path.join(HERE, fname)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        head, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(head)}
        first = next(r for r in rows[2:] if match in r[col['Kernel Name']])   # first launch of the named kernel in the capture

        def val(name):
            v, u = fnum(first[col[name]]), units[col[name]]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return v * scale
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        t, tu = fnum(first[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]]
        us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[tu]
        out.append({"kernel": kernel, "workload": workload, "dram_bytes": rd + wr, "dram_read_bytes": rd,
                    "dram_write_bytes": wr, "us_under_ncu": us,
                    "source": f"ncu --set full --clock-control none, profiles/{fname} ({first[col['Kernel Name']][:60]}...)"})
    json.dump(out, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    for r in out:
        print(f"{r['kernel']:32s} {r['workload']:10s} {r['dram_bytes'] / 1e6:8.1f} MB  {r['us_under_ncu']:7.1f} us")


if __name__ == "__main__":
    main()
