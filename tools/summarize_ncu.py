"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md "<command>"
  python tools/summarize_ncu.py full     gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_gemm_full.md
  python tools/summarize_ncu.py rawcsv   gpurun_out/full_raw.csv profiles/r2_kernels_full.md "<command>"     (raw page exported on the box:
                                         one column per kernel KIND = (name, grid size), the median-duration launch of the kind)
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1000 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000)
        agg[name][0] += 1
        agg[name][1] += us
        seq.append((name, us))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"command: `{cmd}`\n\ncaptured launches: {len(seq)}, total {tot:.1f} us\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write("\nfirst 60 launches in order (name, us):\n\n```\n")
        for n, u in seq[:60]:
            f.write(f"{n[:60]:60s} {u:9.1f}\n")
        f.write("```\n")
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({src.split('/')[-1]}), one column per captured launch\n\n")
        names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]) for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"#{i} {n[:40]}" for i, n in enumerate(names)) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for w in WANT:
            if w in idx:
                f.write(f"| {w} | {units[idx[w]]} | " + " | ".join(r[idx[w]] for r in rows[2:]) + " |\n")
    print("wrote", dst)
    # dominant-kernel DRAM traffic for bench.py's roofline.traffic
    try:
        import json, os
        def num(r, key):
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]].lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
        tot = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in rows[2:]]
        out = {"kernel": names[0], "dram_bytes_per_launch": sum(tot) / len(tot), "source": os.path.basename(dst), "launches": len(tot)}
        with open(os.path.join(os.path.dirname(dst), "ncu_traffic.json"), "w") as fh:
            json.dump(out, fh)
        print("wrote ncu_traffic.json", out)
    except Exception as exc:   # pragma: no cover
        print("traffic summary skipped:", exc)


WANT2 = WANT[:8] + ["smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
                    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
                    "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def rawcsv(src, dst, cmd, persist_steps=1):
    import json, os
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kinds = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("rp::", "")
        kinds.setdefault((name[:70], r[idx["launch__grid_size"]]), []).append(r)

    def val(r, key):
        return float(r[idx[key]].replace(",", "")) if r[idx[key]] not in ("", "n/a") else float("nan")

    def scale(key):
        u = units[idx[key]].lower()
        return 1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0
    cols, traffic = [], {}
    for (name, grid), rs in kinds.items():
        rs = sorted(rs, key=lambda r: val(r, "gpu__time_duration.sum"))
        cols.append((f"{name} grid={grid} (n={len(rs)})", rs[len(rs) // 2]))
        traffic[f"{name} grid={grid}"] = sum(val(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum") +
                                             val(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum") for r in rs) / len(rs)
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, one column per kernel kind (name, grid), the median-duration launch of the kind; every cell carries its own unit\n\n")
        f.write(f"command: `{cmd}`\n\n")
        f.write("| metric | " + " | ".join(c[0] for c in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        for w in WANT2:
            if w in idx:
                f.write(f"| {w} | " + " | ".join(f"{c[1][idx[w]]} {units[idx[w]]}" for c in cols) + " |\n")
    top = max(traffic, key=lambda k: ("fwd" in k, traffic[k]))
    per_step = traffic[top] / (persist_steps if "persist" in top else 1)
    with open(os.path.join(os.path.dirname(dst), "ncu_traffic.json"), "w") as fh:
        json.dump({"kernel": top, "dram_bytes_per_launch": per_step, "per_kernel": traffic, "source": os.path.basename(dst),
                   "note": "dominant kernel: bytes per Euler step (the persistent launch of the capture covered %d steps); per_kernel: bytes per launch" % persist_steps},
                  fh, indent=1)
    print("wrote", dst, "and ncu_traffic.json")


if __name__ == "__main__":
    if sys.argv[1] == "rawcsv":
        rawcsv(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "", int(sys.argv[5]) if len(sys.argv) > 5 else 1)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
