"""Turn gpurun_out ncu artefacts into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_kernels.md [cfg]
"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum",
    "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("nafae::<unnamed>::", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] in ("msecond", "ms"):
            v *= 1e3
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        fh.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        fh.write("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            fh.write("| `%s` | %d | %.2f | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print(open(dst).read())


def full(src, dst, cfg):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(dst, "w") as fh:
        fh.write("# ncu --set full summary (%s)\n" % os.path.basename(src))
        seen = set()
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            if name in seen:
                continue
            seen.add(name)
            fh.write("\n## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % name)
            for k in KEYS:
                if k in d:
                    fh.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            try:
                mb = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
                traffic[name.split("<")[0] + "_dram_bytes"] = int(mb * scale)
            except Exception:
                pass
    tp = os.path.join(os.path.dirname(dst), "ncu_traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    cur.setdefault(cfg, {}).update(traffic)
    if any(k.startswith("align_pool_fwd_slab") for k in traffic):
        # bench.py reports roofline.traffic only while the kernel source is the one that was profiled
        import hashlib
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        src_path = os.path.join(root, "nafae_b200", "csrc", "roi_align.cu")
        cur[cfg]["align_pool_fwd_slab_source_sha"] = hashlib.sha256(open(src_path, "rb").read()).hexdigest()[:16]
    json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "cfg2")
