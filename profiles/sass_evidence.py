"""Static SASS instruction evidence per kernel (cuobjdump -sass of the built library) -> profiles/r2_sass_evidence.txt."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "successiveconvexification_b200", "libscvx_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
out = ["SASS evidence (cuobjdump -sass libscvx_b200.so, sm_100a), static instruction counts per kernel"]
keys = ['UBLKCP', 'UBLKPF', 'USETMAXREG', 'SYNCS', 'DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDG', 'STG', 'RED', 'ATOMG', 'SHFL', 'MUFU', 'LDL', 'STL']
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function :|\Z)", txt, re.S):
    name, body = m.group(1), m.group(2)
    ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body))
    short = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f_]+?_cu_[0-9a-f]+\d*", "", name)
    agg = collections.Counter()
    for k, v in ops.items():
        base = k.split('.')[0]
        if base in keys:
            agg[base] += v
        if k.startswith(('UBLKCP', 'UBLKPF', 'USETMAXREG', 'SYNCS')) or k in ('LDS.128', 'STS.128', 'LDG.E.64.CONSTANT'):
            agg[k] += v
    out.append(f"\n{short}\n  total {sum(ops.values())}  " + "  ".join(f"{k}={v}" for k, v in sorted(agg.items())))
open(os.path.join(ROOT, "profiles", "r2_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
