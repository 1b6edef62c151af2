"""Write profiles/r2_ncu_traffic.json (what bench.py's roofline.traffic reads) from an ncu --set full
report of one DyT block (profiles/prof_step.py --layers 1), stamped with the sha256 of the kernel
sources it was taken with:  python profiles/make_traffic_record.py gpurun_out/prof_r2m.ncu-rep r2m"""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
# launch order of one block (prof_step.py --layers 1): LN1, qkv, attention, proj, down, dispatcher, fc1, fc2, merge
names = [None, "gemm qkv [T,768]x[2304,768]", "attention 12 heads x 197", "gemm proj + residual",
         "gemm adapter down + ReLU", "dispatcher (score+gate+compact+LN2 pack)", "gemm fc1 + GELU (kept rows)",
         "gemm fc2 (kept rows)", "adapter up + scatter-merge + next LN1 (fused)"]
traffic = {}
for n, r in enumerate(rows[2:]):
    if n < len(names) and names[n]:
        f = lambda k: float(r[ix[k]].replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = f("dram__bytes_read.sum") * scale[rows[1][ix["dram__bytes_read.sum"]]]
        wr = f("dram__bytes_write.sum") * scale[rows[1][ix["dram__bytes_write.sum"]]]
        traffic[names[n]] = rd + wr
files = ["dynamic-tuning_b200/csrc/gemm_tn.cuh", "dynamic-tuning_b200/csrc/gemm_tn.cu",
         "dynamic-tuning_b200/csrc/ptx.cuh", "dynamic-tuning_b200/csrc/gelu.cuh"]
h = hashlib.sha256()
for f in files:
    h.update(open(os.path.join(ROOT, f), "rb").read())
rec = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, "
               f"one DyT block at B=256 (profiles/{tag}_ncu_full_one_layer.md)",
       "capture": f"{rep} (not committed); summary profiles/{tag}_ncu_full_one_layer.md",
       "source_files": files, "source_sha256": h.hexdigest(), "dram_bytes_per_launch": traffic}
json.dump(rec, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
