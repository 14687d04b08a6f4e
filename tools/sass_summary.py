"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): UTCIMMA / UTCHMMA
(tcgen05.mma int8 / f16), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA loads), SYNCS (mbarrier), USETMAXREG, VOTE.
    python tools/sass_summary.py > profiles/sass_summary.txt      (runs cuobjdump -sass on the built libsnnqp.so)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "snnquantprune_b200", "libsnnqp.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
keys = ["UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG", "VOTE", "FFMA2", "IMMA", "HMMA"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
  m = re.search(r"Function : (\S+)", line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    continue
  m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
  if m and cur:
    op = m.group(1)
    counts[cur]["_n"] += 1
    for k in keys:
      if op == k or op.startswith(k + "."):
        counts[cur][k] += 1
        total[k] += 1
def demangle(n):
  try:
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    m = re.search(r"(snnqp::[\w:]+(?:<[^(]*>)?)", d)
    return (m.group(1) if m else d)[:110]
  except Exception:
    return n[:90]
print(f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a); columns: instructions, then mnemonic counts")
print("# " + " ".join(f"{k:>9s}" for k in ["instrs"] + keys) + "  kernel")
for name, c in counts.items():
  if not any(c[k] for k in keys[:6]):
    continue
  print("  " + " ".join(f"{c[k]:9d}" for k in ["_n"] + keys) + "  " + demangle(name))
print("# total " + " ".join(f"{k}={total[k]}" for k in keys))
