"""Aggregates an ncu source-page CSV (SASS level) by CUDA source line using nvdisasm -g line info.

usage: python profiles/hotlines.py <report.ncu-rep | source-page.csv> <kernel regex> <cubin> [top_n] [cubin symbol regex] [outer]
With `outer` as the sixth argument samples are attributed to the outermost source line of the inline chain
(the statement of the kernel body), i.e. to the phase of the kernel instead of the helper it inlines.
The SASS instruction order of the ncu page and of nvdisasm agree, so instruction k of the kernel
is mapped to the `//## File "...", line N` annotation that precedes it in the nvdisasm listing.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kregex, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    cregex = sys.argv[5] if len(sys.argv) > 5 else kregex  # symbol regex inside the cubin
    outer = len(sys.argv) > 6 and sys.argv[6] == "outer"
    if rep.endswith(".csv"):   # a `--page source --csv` export made on the GPU box (profiles/capture.sh)
        src = open(rep).read()
    else:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex],
                             capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # may contain several launches: keep the first kernel block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    dis = subprocess.run(["nvdisasm", "-gi" if outer else "-g", "-c", cubin], capture_output=True, text=True).stdout
    # find the function section matching the kernel
    lines = dis.splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and re.search(cregex, l):
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found in cubin")
    cur = ("?", 0)
    mapping = []
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            if mapping:
                break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if outer:
                allm = re.findall(r'"([^"]+)", line (\d+)', l)
                m_file, m_line = allm[-1]
                cur = (m_file.split("/")[-1], int(m_line))
            else:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            mapping.append(cur)
    n = min(len(mapping), len(body))
    agg = defaultdict(lambda: [0, 0])
    for k in range(n):
        r = body[k]
        agg[mapping[k]][0] += int(r[col["# Samples"]] or 0)
        agg[mapping[k]][1] += int(r[col["Instructions Executed"]] or 0)
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    print("instructions in page %d, in disasm %d; total samples %d, warp-instructions %d" % (len(body), len(mapping), tot_s, tot_i))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s:%-5d samples %6.2f%%  instr %6.2f%%" % (key[0], key[1], 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i))
    # stall reasons overall
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
    s = sum(tot.values()) or 1
    print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / s) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
