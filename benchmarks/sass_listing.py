"""python benchmarks/sass_listing.py <mangled-substring> <title> <out> [interesting-regex]: mnemonic histogram + selected lines."""
import re, subprocess, sys, collections
sub, title, out = sys.argv[1], sys.argv[2], sys.argv[3]
pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 else re.compile(r"LDGSTS|SYNCS|ARRIVES|UBLKCP|ATOMG|RED\.|BAR\.SYNC")
txt = subprocess.run(["cuobjdump", "-sass", "advmix_b200/libadvmix_b200.so"], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
body = [b for b in blocks if sub in b.split("\n")[0]][0]
lines = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
hist = collections.Counter()
for l in lines:
    t = l.split("*/", 1)[1].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    hist[op.rstrip(";")] += 1
with open(out, "w") as f:
    f.write("# SASS of %s (sm_100a), cuobjdump -sass advmix_b200/libadvmix_b200.so\n# %d instructions; mnemonic histogram:\n" % (title, len(lines)))
    for op, n in hist.most_common(40):
        f.write("%7d %s\n" % (n, op))
    f.write("\n# instructions of interest:\n")
    sel = [l for l in lines if pat.search(l)]
    for l in sel[:60]:
        f.write(l.rstrip() + "\n")
print(out, len(lines))
