"""Per-source-line executed instructions and stall samples of kpp_step_kernel from an ncu report.

  ncu --section SourceCounters --import-source on ... -o rep      (on the GPU box, tools/ncu_src.sh)
  python tools/src_lines.py gpurun_out/rep.ncu-rep [object.o] [top]

Joins the SASS page of the report (per-instruction counters) with nvdisasm's line table of the
object that was profiled, because the CUDA-source page of the CSV export carries no counters."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "mckpp_f90_b200/build/kpp_kernels_strict.o")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
kern = os.environ.get("KERNEL", "kpp_step_kernel_strict")
src = os.path.join(ROOT, "mckpp_f90_b200/csrc/kpp_kernels.cu")

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lm, line, on = {}, None, False
for l in dis.split("\n"):
    if l.startswith(".text.") and kern in l:
        on = True
        continue
    if on and l.startswith("//-----") and lm:
        break
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        lm[int(m.group(1), 16)] = line

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, isamp, iex = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
reasons = [(n[6:], i) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
why = collections.defaultdict(collections.Counter)
inst, samp, base = collections.Counter(), collections.Counter(), None
for r in rows[hi + 1:]:
    try:
        ad = int(r[ia], 16)
    except (ValueError, IndexError):
        continue
    base = ad if base is None else base
    ln = lm.get(ad - base)
    inst[ln] += int(r[iex] or 0)
    samp[ln] += int(r[isamp] or 0)
    for name, i in reasons:
        v = int(r[i] or 0) if i < len(r) else 0
        if v:
            why[ln][name] += v
text = open(src).read().split("\n")
print(f"total: {sum(inst.values())/1e6:.1f} M warp instructions, {sum(samp.values())} samples")
tot_why = collections.Counter()
for c in why.values():
    tot_why.update(c)
print("stall reasons (all samples): " + ", ".join(f"{k} {v}" for k, v in tot_why.most_common(8)))
for ln, v in inst.most_common(top):
    top2 = " ".join(f"{k}:{n}" for k, n in why[ln].most_common(3))
    print(f"L{ln}: {v/1e6:7.1f} M inst {samp[ln]:6d} samp [{top2}] | {text[ln-1].strip()[:90] if ln else ''}")
