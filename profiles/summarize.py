"""Turns an ncu report into the committed evidence: a per-kernel CSV/markdown table and
profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).

usage: python profiles/summarize.py <report.ncu-rep | raw-page.csv> <tag>
"""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("smsp__inst_executed.sum", "warp_inst_M"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefront_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
]


def csrc_sha1(kernel=None):
    """SHA-1 over the sources a kernel is compiled from: the .cu file that defines it plus every header of
    world-class_b200/csrc (sorted by name).  Without a kernel name (or when no file defines it): all sources."""
    import hashlib
    import re
    d = os.path.join(HERE, "..", "world-class_b200", "csrc")
    names = sorted(os.listdir(d))
    files = names
    if kernel:
        base = re.sub(r"<.*", "", kernel)
        owners = [n for n in names if n.endswith(".cu") and re.search(r"\b%s\b" % re.escape(base), open(os.path.join(d, n)).read())]
        # (the file that DEFINES it: a __global__ definition, not a mention in a comment of another file)
        owners = [n for n in owners if re.search(r"__global__[^;{]*\b%s\s*\(" % re.escape(base), open(os.path.join(d, n)).read(), re.S)] or owners
        if owners:
            files = sorted(set(owners) | {n for n in names if n.endswith((".cuh", ".h"))})
    h = hashlib.sha1()
    for name in files:
        h.update(name.encode())
        h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # a `--page raw --csv` export made on the GPU box (profiles/capture.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out_rows = []
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        rec = {"kernel": name}
        for m, short in METRICS:
            if m in idx:
                v = r[idx[m]].replace(",", "")
                try:
                    v = float(v)
                    u = units[idx[m]]
                    if short == "time_us":
                        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
                    if short.endswith("_MB"):
                        v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    if short == "warp_inst_M":
                        v = v / 1e6
                except ValueError:
                    pass
                rec[short] = v
        out_rows.append(rec)
        base = name.split("<")[0]
        if "dram_read_MB" in rec and "dram_write_MB" in rec:
            traffic.setdefault(base, []).append((rec["dram_read_MB"] + rec["dram_write_MB"]) * 1e6)
    cols = ["kernel"] + [s for _, s in METRICS]
    with open(os.path.join(HERE, "ncu_%s.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(cols)
        for rec in out_rows:
            w.writerow([rec.get(c, "") if not isinstance(rec.get(c), float) else "%.4g" % rec[c] for c in cols])
    # merged into the existing file (a capture usually covers a few kernels), stamped with a hash of the kernel
    # sources: bench.py reports `traffic_stale` when the sources have changed since
    path = os.path.join(HERE, "ncu_traffic.json")
    merged = {}
    if os.path.exists(path):
        try:
            merged = json.load(open(path))
        except Exception:
            merged = {}
    for k, v in traffic.items():
        merged[k] = max(v)
        merged.setdefault("_csrc_sha1_by_kernel", {})[k] = csrc_sha1(k)
    with open(path, "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)
    for rec in out_rows:
        print(" | ".join("%s" % (("%.4g" % rec[c]) if isinstance(rec.get(c), float) else rec.get(c, "")) for c in cols))


if __name__ == "__main__":
    main()
