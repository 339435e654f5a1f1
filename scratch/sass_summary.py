"""cuobjdump -sass opcode histogram per kernel of scanpaths_b200/build/*.o -> profiles/r02_sass_summary.txt."""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPECIAL = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "LDGSTS", "SHFL", "REDUX")
out = ["# SASS opcode summary of the sm_100a kernels (cuobjdump -sass on scanpaths_b200/build/*.o; nvcc 12.9, -arch compute_100a/sm_100a)",
       "# per kernel: total instructions, then the Blackwell-specific opcodes (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store,",
       "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async) and the most frequent opcodes", ""]
for obj in sorted(glob.glob(os.path.join(ROOT, "scanpaths_b200", "build", "*.o"))):
    out.append("## " + os.path.relpath(obj, ROOT))
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, cnt = None, None
    def flush():
        if name and cnt:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = re.sub(r"\((int|bool)\)", "", dem)
            dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("spb::", "")
            sp = " ".join("%s=%d" % (k, v) for k, v in sorted((k, sum(c for o, c in cnt.items() if o.startswith(k))) for k in SPECIAL) if v)
            top = " ".join("%s=%d" % kv for kv in cnt.most_common(8))
            out.append("%-50s %5d instr | %s | top: %s" % (dem, sum(cnt.values()), sp or "-", top))
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            name, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cnt is not None:
            cnt[m.group(1).split(".")[0]] += 1
    flush()
    out.append("")
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out))
print("\n".join(l for l in out if "wino" in l or "conv_gemm" in l))
