"""Writes profiles/r02_sass_final.txt: mnemonic counts and excerpts of `cuobjdump -sass libifb200.so` that show TMA,
tcgen05 (incl. the cta_group::2 forms), TMEM, mbarrier, bulk copies and the L2 prefetch in the shipped library."""
import collections
import re
import subprocess
import sys

so = "isolation-forest_b200/libifb200.so"
sass = subprocess.check_output(["cuobjdump", "-sass", so]).decode()
cnt = collections.Counter()
for m in re.finditer(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass, re.M):
    op = m.group(1)
    if re.match(r"(UTMA|UTC|LDTM|STTM|SYNCS|UBLKCP|UTMAPF|FMNMX3)", op):
        cnt[op] += 1
out = ["# SASS evidence of the final round-2 library (cuobjdump -sass libifb200.so, sm_100a; tools/sass_evidence.py)\n",
       "## mnemonic counts (TMA, tcgen05 incl. the 2-CTA forms, TMEM, mbarrier, bulk copies, L2 prefetch, 3-input min)\n"]
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]):
    out.append(f"    {v:5d} {k}\n")


def excerpt(fun_pat, keys, title, ctx=0, maxn=14):
    out.append(f"\n## {title}\n")
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0]
        if not re.search(fun_pat, name):
            continue
        lines = blk.split("\n")
        n = 0
        for i, l in enumerate(lines):
            if any(k in l for k in keys) and "/*" in l:
                for j in range(max(0, i - ctx), min(len(lines), i + ctx + 1)):
                    if re.search(r"/\*[0-9a-f]{4,6}\*/\s+\S", lines[j]) and not lines[j].strip().startswith("/* 0x"):
                        out.append("    " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", lines[j]).rstrip() + "\n")
                out.append("    ...\n")
                n += 1
                if n >= maxn:
                    break
        out.append(f"    (function {name[:120]})\n")
        break


excerpt(r"score_ext_tc_kernelILb0ELi2ELi32ELi1", ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "FMNMX3"],
        "score_ext_tc_kernel<false, 2, 32, 1> (default for narrow hyperplanes): TMA loads (multicast), tcgen05.mma, "
        "tcgen05.commit, tcgen05.ld, the drain's 3-input min", 0, 16)
excerpt(r"score_ext_tc_kernelILb0ELi2ELi32ELi2", ["UTCHMMA", "UTCBAR", "UTMALDG", "SYNCS.ARRIVE"],
        "score_ext_tc_kernel<false, 2, 32, 2> (IFB_TC_CG=2): the cta_group::2 pair -- 2-CTA MMA, 2-CTA TMA loads, 2-CTA "
        "multicast commits", 0, 14)
excerpt(r"score_std_kernelILi256ELi16ELb1ELb0ELi6ELi1", ["UTMAPF", "UTMALDG"],
        "score_std_kernel<256, 16, TMA, no depth, 6 deep levels, 1 stage>: L2 prefetch of the next tile + the tile's TMA loads",
        0, 6)
excerpt(r"score_std_rank_kernelILi6", ["UBLKCP", "SYNCS.ARRIVE", "FSETP.GE"],
        "score_std_rank_kernel<6> (opt-in): bulk copies of the feature columns; a visit = LDS, LOP3, LOP3, LDS, FSETP, @P LOP3",
        2, 4)
open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sass_final.txt", "w").writelines(out)
print("".join(out[:24]))
