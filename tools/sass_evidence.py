"""profiles/*_sass_evidence.txt: what the built libvxp.so contains (cuobjdump -sass, here, no GPU needed): target
architectures, per-kernel instruction counts of the mnemonics that matter for this path, and short excerpts around the
first bulk-TMA copy / mbarrier wait / 32-byte vertex store.   usage: python tools/sass_evidence.py <out.txt>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "voxelphile_b200", "libvxp.so")
out = open(sys.argv[1], "w")
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
out.write(f"libvxp.so at commit {commit}: cuobjdump -lelf\n{elf}\n")
MN = ["UBLKCP", "SYNCS", "STG.E.ENL2.256", "STG.E.64", "STG.E.128", "LDG", "LDS", "STS", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "BREV", "FLO", "POPC",
      "IDP", "VIMNMX", "PRMT", "BAR", "ACQBULK", "CCTL", "UTMALDG", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA", "IMMA", "LDL", "STL"]
fn = None; per = collections.OrderedDict(); lines = {}
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[fn] = collections.Counter(); lines[fn] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if fn and m:
        ins = m.group(2).strip()
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        per[fn]["total"] += 1
        for k in MN:
            if op.startswith(k): per[fn][k] += 1
        lines[fn].append(f"/*{m.group(1)}*/ {m.group(2).strip()}")
out.write("instruction counts per kernel (static SASS; tensor-core / TMEM / tensor-map mnemonics are expected to be 0: nothing\n"
          "on this path is a contraction, and the tile copy is the 1-D bulk form)\n\n")
hdr = ["kernel", "total"] + MN
out.write(" | ".join(hdr) + "\n")
for f, c in per.items():
    out.write(" | ".join([f] + [str(c.get(k, 0)) for k in hdr[1:]]) + "\n")
for f in per:
    if "mesh_kernel" not in f: continue
    out.write(f"\n== {f}: first bulk copy, first mbarrier try-wait, first 32-byte vertex store (3 lines of context)\n")
    for pat in ("UBLKCP", "SYNCS.PHASECHK", "STG.E.ENL2.256"):
        for i, l in enumerate(lines[f]):
            if pat in l:
                out.write("\n".join(lines[f][max(0, i - 3): i + 4]) + "\n   ...\n")
                break
out.close()
