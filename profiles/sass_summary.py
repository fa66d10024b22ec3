"""SASS instruction mix of the hot kernels of libworldb200.so (cuobjdump -sass): what the compiled code consists of --
fp64 arithmetic (DADD / DMUL / DFMA), fp64 tensor-pipe MMAs (DMMA), shared / global memory instructions, barriers,
shuffles -- per kernel.  usage: python profiles/sass_summary.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "world-class_b200", "libworldb200.so")
HOT = ["d4c_body_kernelILi12", "refine_mma", "channel_kernelILi13", "response_kernelILi11", "ct_frame_kernelILi11", "lt_frame_kernelILi13",
       "rng_fill", "interval_kernel", "candidate_kernel", "remove_kernel", "harvest_tail"]
GROUPS = [("fp64 add/mul/fma", r"^(DADD|DMUL|DFMA)"), ("fp64 tensor MMA (DMMA)", r"^DMMA"), ("fp64 other (DSETP, MUFU.RCP64H, F2F/F2I/I2F .F64)", r"^(DSETP|DMNMX|MUFU|F2I|I2F|F2F)"),
          ("shared-memory load/store", r"^(LDS|STS)"), ("global load/store", r"^(LDG|STG|LD\.|ST\.)"), ("shared atomics", r"^ATOMS"),
          ("block barrier (BAR)", r"^BAR"), ("warp sync / shuffle / vote / redux", r"^(WARPSYNC|SHFL|VOTE|REDUX|MATCH)"),
          ("bulk copy / mbarrier (UBLKCP, SYNCS)", r"^(UBLKCP|SYNCS|UTMALDG)"), ("integer / logic / move", r"^(IMAD|IADD|LOP|SHF|LEA|MOV|SEL|ISETP|PRMT|BREV|VIADD|IABS|IMNMX|UMOV|ULDC|S2R|CS2R|UIADD|USHF|ULOP|UISETP|USEL|UIMAD|ULEA|R2UR|PLOP|P2R|R2P|FLO|POPC|I2I)"),
          ("branch / control", r"^(BRA|EXIT|BSSY|BSYNC|CALL|RET|NOP|YIELD|BREAK|WARPSYNC|DEPBAR|ERRBAR|MEMBAR|LDC|LDL|STL)")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
for key in HOT:
    for fn, c in counts.items():
        if key in fn:
            total = sum(c.values())
            print("%s  (%d instructions)" % (fn, total))
            rest = collections.Counter(c)
            for name, pat in GROUPS:
                n = sum(v for k, v in c.items() if re.match(pat, k))
                for k in list(rest):
                    if re.match(pat, k):
                        del rest[k]
                if n:
                    print("    %-55s %6d  %5.1f %%" % (name, n, 100.0 * n / total))
            if rest:
                print("    %-55s %6d  %5.1f %%   (%s)" % ("other", sum(rest.values()), 100.0 * sum(rest.values()) / total, ", ".join(k for k, _ in rest.most_common(6))))
            print()
            break
