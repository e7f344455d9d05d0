"""Turn gpurun_out/ ncu artefacts into small text summaries under profiles/ (tracked).

  python tools/summarize_profiles.py launches gpurun_out/launches5.csv profiles/r01_step_launches.md
  python tools/summarize_profiles.py ncu gpurun_out/prof_gemm3.ncu-rep profiles/r01_gemm_fc1_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def launches(src, dst):
    with open(src) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    half = rows[len(rows) // 2:]          # tools/profile_step.py runs 2 steps: keep the second (steady state)
    agg = collections.OrderedDict()
    for r in half:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as o:
        o.write(f"# ncu launch list, one training step (second of two), {len(half)} launches, sum {tot/1e6:.3f} ms\n")
        o.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_step.py 2\n")
        o.write("# (per-launch times are cold-cache and serialised: compare SHARES)\n\n")
        o.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| `{k[:90]}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f}% |\n")
    print("wrote", dst)


def launches_dram(src, dst, traffic_json=None):
    """Launch list captured with gpu__time_duration.sum + dram__bytes_read.sum + dram__bytes_write.sum (one CSV row per
    launch per metric). Also writes profiles/gemm_traffic.json = mean DRAM bytes per ViT GEMM launch, which bench.py
    reports as roofline.traffic (it cannot measure DRAM traffic itself)."""
    import json
    with open(src) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    by_id = collections.OrderedDict()
    for r in rows:
        d = by_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"])
    ids = list(by_id)
    # tools/profile_step.py runs 2 steps: keep the last one. A step starts with the patch re-tiling kernel (one launch per
    # step); halving the list would mis-split, the first step carries one-time launches (weight down-casts, ...)
    starts = [i for i in ids if "patchify_kernel" in by_id[i]["name"]]
    first = starts[-1] if starts else ids[len(ids) // 2]
    half = [by_id[i] for i in ids if i >= first]
    agg = collections.OrderedDict()
    for d in half:
        k = d["name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as o:
        o.write(f"# ncu launch list of one training step (second of two): time + DRAM traffic per kernel family, {len(half)} launches, sum {tot/1e6:.3f} ms\n")
        o.write("# command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/profile_step.py 2\n")
        o.write("# (per-launch times are cold-cache and serialised: compare SHARES with the CUDA-graph step bench.py times)\n\n")
        o.write("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | avg GB/s |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| `{k[:90]}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f}% | {v[2]/1e6:.1f} | {v[3]/1e6:.1f} | {(v[2]+v[3])/max(v[1],1):.0f} |\n")
        # the big ViT GEMMs: tcgen05 GEMM launches on a full persistent grid (CTA pairs, 148 CTAs)
        # (the generic / plain-store MN/MN-major instances are left out: those are the AVT-h weight gradients over 80 contraction rows)
        big = [d for d in half if "gemm_bf16_kernel<256, 2" in d["name"]
               and not any(t in d["name"] for t in ("<256, 2, 1, 1, 0>", "<256, 2, 1, 1, 1>", "<256, 2, 1, 1, 5>"))]
        if big:
            b = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in big) / len(big)
            t = sum(d.get("gpu__time_duration.sum", 0.0) for d in big)
            o.write(f"\nViT GEMMs (256-wide CTA-pair launches on the full grid): {len(big)} launches, {t/1e3:.1f} us, mean DRAM traffic {b/1e6:.1f} MB per launch\n")
            if traffic_json:
                with open(traffic_json, "w") as j:
                    # keyed by "<model>:<frames>:<clips per GPU>" (bench.py looks its own workload up); profile_step.py runs cfg2
                    json.dump({"vit_base_patch16_224:10:8": {
                        "bytes_per_launch": b, "launches": len(big), "workload": "cfg2",
                        "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the {len(big)} ViT GEMM launches of one step ({dst})"}},
                        j, indent=1)
    print("wrote", dst)


def ncu(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as o:
        o.write(f"# ncu --set full --clock-control none --import-source on  ({src})\n")
        for d in data:
            o.write(f"\n## {d[hdr.index('Kernel Name')][:110]}  grid {d[hdr.index('Grid Size')]} block {d[hdr.index('Block Size')]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    o.write(f"{k:75s} {d[i]:>16s} {units[i]}\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "launches_dram": launches_dram, "ncu": ncu}[sys.argv[1]](*sys.argv[2:])
