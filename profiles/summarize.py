"""Turn the raw capture of one round (gpurun_out/<prefix>_*) into the committed summaries under profiles/ (run here, no GPU).
    python profiles/summarize.py r2"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
shutil.copy(os.path.join(G, f"{R}_launches.csv"), os.path.join(P, f"{R}_launches.csv"))
for src, dst in ((f"{R}_bench.json", f"{R}_bench_1gpu.json"), (f"{R}_bench_reference.json", f"{R}_bench_reference.json"),
                 (f"{R}_bench_textbook.json", f"{R}_bench_1gpu_textbook.json"),
                 (f"{R}_bench_sigma_near_1.json", f"{R}_bench_1gpu_sigma_near_1.json")):
    if os.path.exists(os.path.join(G, src)):
        txt = open(os.path.join(G, src)).read()
        open(os.path.join(P, dst), "w").write([l for l in txt.splitlines() if l.startswith("{")][-1] + "\n")
kernels = [k for k in ("tangent_kernel", "stage_value_kernel", "compact_pack_kernel") if os.path.exists(os.path.join(G, f"{R}_{k}.ncu-rep"))]
for k in kernels:
    out = subprocess.run([sys.executable, os.path.join(P, "ncu_summary.py"), os.path.join(G, f"{R}_{k}.ncu-rep")],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{R}_{k}.txt"), "w").write(out)
rows = [r for r in csv.reader(open(os.path.join(P, f"{R}_launches.csv"))) if r and not r[0].startswith("==")]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ik].split("(")[0].split("::")[-1].split("<")[0]; tot[name] += float(r[iv].replace(",", "")); n[name] += 1
s = sum(tot.values())
shares = {k: {"launches": n[k], "total_us": round(v / 1e3, 1), "share": round(v / s, 3)} for k, v in tot.items()}
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
traffic = {}
for k in kernels:
    out = subprocess.run(["ncu", "-i", os.path.join(G, f"{R}_{k}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines())); h, u, v = rr[0], rr[1], rr[2]
    def get(name):
        i = h.index(name); return float(v[i].replace(",", "")) * mult.get(u[i], 1)
    tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[u[h.index("gpu__time_duration.sum")]]
    traffic[k] = {"dram_read": get("dram__bytes_read.sum"), "dram_write": get("dram__bytes_write.sum"),
                  "duration_us": float(v[h.index("gpu__time_duration.sum")].replace(",", "")) * tu,
                  "grid": int(float(v[h.index("launch__grid_size")].replace(",", "")))}
t = traffic["tangent_kernel"]
# the captured launches are the same chunk of the same step (-s 4 -c 1 of either kernel): its intervals = value-kernel grid x 128
n_int = traffic["stage_value_kernel"]["grid"] * 128
json.dump({"kernel": "tangent_kernel", "dram_bytes_per_launch": t["dram_read"] + t["dram_write"], "dram_read": t["dram_read"],
           "dram_write": t["dram_write"], "intervals_per_launch": n_int,
           "dram_bytes_per_interval": (t["dram_read"] + t["dram_write"]) / n_int, "per_kernel": traffic, "device_time_shares": shares,
           "note": f"one chunk of the capture run ({n_int} intervals); ncu --set full, see {R}_tangent_kernel.txt; shares from "
                   f"{R}_launches.csv (cold-cache, serialised launches: compare shares, not absolutes)"},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(shares, indent=1))
