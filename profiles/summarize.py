"""Turn the raw capture (gpurun_out/r1_*) into the committed summaries under profiles/ (run here, no GPU)."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "r1_launches.csv"), os.path.join(P, "r1_launches.csv"))
for src, dst in (("r1_bench.json", "r1_bench_1gpu.json"), ("r1_bench_reference.json", "r1_bench_reference.json")):
    txt = open(os.path.join(G, src)).read()
    open(os.path.join(P, dst), "w").write([l for l in txt.splitlines() if l.startswith("{")][-1] + "\n")
for k in ("tangent_kernel", "stage_value_kernel"):
    out = subprocess.run([sys.executable, os.path.join(P, "ncu_summary.py"), os.path.join(G, f"r1_{k}.ncu-rep")],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"r1_{k}.txt"), "w").write(out)
rows = [r for r in csv.reader(open(os.path.join(P, "r1_launches.csv"))) if r and not r[0].startswith("==")]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ik].split("(")[0].split("::")[-1]; tot[name] += float(r[iv].replace(",", "")); n[name] += 1
s = sum(tot.values())
shares = {k: {"launches": n[k], "total_us": round(v / 1e3, 1), "share": round(v / s, 3)} for k, v in tot.items()}
out = subprocess.run(["ncu", "-i", os.path.join(G, "r1_tangent_kernel.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines())); h, u, v = rr[0], rr[1], rr[2]
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
def get(name):
    i = h.index(name); return float(v[i].replace(",", "")) , u[i]
rd, ru = get("dram__bytes_read.sum"); wr, wu = get("dram__bytes_write.sum")
json.dump({"kernel": "tangent_kernel", "dram_bytes_per_launch": rd * mult[ru] + wr * mult[wu], "dram_read": rd * mult[ru],
           "dram_write": wr * mult[wu], "intervals_per_launch": 113664, "device_time_shares": shares,
           "note": "one full chunk (148 SMs x 768 intervals); ncu --set full, see r1_tangent_kernel.txt; "
                   "shares from r1_launches.csv (cold-cache, serialised launches: compare shares, not absolutes)"},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
d = json.load(open(os.path.join(P, "r1_bench_1gpu.json")))
print(shares)
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "peak", d["roofline"]["peak"], d["clocks"],
      "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
