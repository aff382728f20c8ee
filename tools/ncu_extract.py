"""Summarise an `ncu --set full` capture of the step kernel and stamp it into profiles/traffic.json.

    python tools/ncu_extract.py gpurun_out/X.ncu-rep "<workload name>" <columns> profiles/r2_ncu_X.csv

Writes the metric summary CSV and adds {kernel_sha, workload, bytes_per_launch, ...} to
profiles/traffic.json; bench.py reports `roofline.traffic` only from a capture whose kernel_sha
matches the kernel source it is running (a stale capture reads as null)."""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload, ncols, out_csv = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.split("\n")))
h, units, v = rows[0], rows[1], rows[2]
col = {n: i for i, n in enumerate(h)}
name = v[col["Kernel Name"]]
vals = {}
with open(out_csv, "w") as f:
    f.write(f"# {name}; {workload}; {ncols} columns; from {os.path.basename(rep)}\nmetric,unit,value\n")
    for w in WANT:
        if w in col:
            vals[w] = v[col[w]]
            f.write(f"{w},{units[col[w]]},{v[col[w]]}\n")


def num(k, scale=1.0):
    return float(vals[k].replace(",", "")) * scale


def to_bytes(k):
    u = units[col[k]].lower()
    return num(k, {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u])


tot = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
u = units[col["gpu__time_duration.sum"]].lower()
t = num("gpu__time_duration.sum", {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}[u])
sha = hashlib.sha256()
for fn in ("kpp_kernels.cu", "kpp_dev.h"):
    sha.update(open(os.path.join(ROOT, "mckpp_f90_b200", "csrc", fn), "rb").read())
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6.65e12
entry = {"kernel": name, "kernel_sha": sha.hexdigest()[:16], "workload": workload, "columns": ncols,
         "source": os.path.relpath(out_csv, ROOT), "bytes_per_launch": tot, "bytes_per_column_step": tot / ncols,
         "kernel_ms": t * 1e3, "dram_frac": tot / t / peak,
         "fp64_pipe_frac": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
         "registers": int(num("launch__registers_per_thread")), "block": int(num("launch__block_size"))}
p = os.path.join(ROOT, "profiles", "traffic.json")
try:
    d = json.load(open(p))
    if "captures" not in d:
        d = {"captures": []}
except Exception:
    d = {"captures": []}
d["captures"] = [e for e in d["captures"] if e["workload"] != workload] + [entry]     # one capture per workload: the latest
json.dump(d, open(p, "w"), indent=1)
print(json.dumps(entry))
