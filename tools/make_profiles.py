"""Post-processes what tools/gpu_profiles.sh left in gpurun_out/ into the tracked evidence under profiles/.
    python tools/make_profiles.py [round tag, default r1]"""
import collections, csv, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r1"

# 1) launch list -> copy + per-kernel summary (launches, ms, DRAM MB)
src = os.path.join(G, f"launches_{R}.csv")
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))[1:]
L = collections.OrderedDict()
for x in rows:
    L.setdefault(int(x[0]), {"name": x[4].split("(")[0]})[x[12]] = float(x[14].replace(",", ""))
summ = collections.OrderedDict()
for v in L.values():
    e = summ.setdefault(v["name"], {"launches": 0, "ms": 0.0, "dram_MB": 0.0})
    e["launches"] += 1
    e["ms"] += v["gpu__time_duration.sum"] / 1e6
    e["dram_MB"] += (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / 1e6
tot = sum(e["ms"] for e in summ.values())
for e in summ.values():
    e["share"] = e["ms"] / tot
shutil.copy(src, os.path.join(P, f"launches_{R}.csv"))
json.dump(summ, open(os.path.join(P, f"launch_summary_{R}.json"), "w"), indent=1)
print(f"launch list: {len(L)} launches, {tot:.3f} ms (cold-cache, serialised)")

# 2) ncu --set full raw pages -> a few columns per launch
KEEP = ["ID", "Kernel Name", "launch__grid_size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for name in (f"igemm_{R}", f"attn_{R}"):
    raw = os.path.join(G, f"{name}_raw.csv")
    if not os.path.exists(raw):
        continue
    r = list(csv.reader(open(raw)))
    hdr, units, body = r[0], r[1], r[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(P, f"{name.split('_')[0]}_ncu_full_{R}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for b in body:
            row = [b[i] for i in idx]
            row[1] = row[1][:60]
            w.writerow(row)
    print(name, len(body), "launches")

# 3) in-pipeline kineto breakdown, elementwise roofline, bench lines
for a, b in (("kineto.txt", f"kineto_step_{R}.txt"), ("kineto_step.json", f"kineto_step_{R}.json"),
             ("elementwise_roofline.jsonl", f"elementwise_roofline_{R}.jsonl"), ("layer_table.txt", f"igemm_layer_table_{R}.txt"),
             ("igemm_roles.txt", f"igemm_roles_{R}.txt"), ("igemm_roles_small.txt", f"igemm_roles_splitk_{R}.txt"),
             ("tma_ingest3.txt", f"micro_tma_issue_{R}.txt"), ("tma_ingest2.txt", f"micro_tma_ingest_{R}.txt"),
             ("pdl_chain.txt", f"micro_pdl_chain_{R}.txt")):
    if os.path.exists(os.path.join(G, a)):
        txt = open(os.path.join(G, a)).read()
        if a == "kineto.txt":
            txt = "\n".join(l for l in txt.splitlines() if "Warn" not in l and "_warn" not in l) + "\n"
        open(os.path.join(P, b), "w").write(txt)
for a, b in (("bench_n1.txt", f"bench_n1_{R}.json"), ("bench_ref.txt", f"bench_reference_{R}.json"), ("bench_n2.txt", f"bench_n2_{R}.json")):
    p = os.path.join(G, a)
    if os.path.exists(p):
        lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if lines:
            open(os.path.join(P, b), "w").write(lines[-1] + "\n")
