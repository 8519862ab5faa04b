"""Turn gpurun_out/*.ncu-rep / launches.csv into small tracked summaries under profiles/ (run in the authoring container)."""
import csv, subprocess, sys, json, collections

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]

def summarize(rep, dst):
    hdr, units, rows = raw(rep)
    idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in rows:
            w.writerow([r[i] for _, i in idx])

def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("void mmc::", "").replace("void ", "")
        if "conv_tc_kernel" in r[kn]:
            name = r[kn].split("(mmc")[0].replace("void mmc::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[val].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct"])
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{us:.1f}", f"{100 * us / tot:.2f}"])

if __name__ == "__main__":
    import os
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    for rep, dst in ((f"gpurun_out/{rnd}_conv_tc_default.ncu-rep", f"profiles/{rnd}_ncu_conv_tc_full.csv"),
                     (f"gpurun_out/{rnd}_entropy.ncu-rep", f"profiles/{rnd}_ncu_entropy_full.csv"),
                     (f"gpurun_out/{rnd}_entropy_cfg2.ncu-rep", f"profiles/{rnd}_ncu_entropy_cfg2_full.csv")):
        if os.path.exists(rep):
            summarize(rep, dst)
    if os.path.exists(f"gpurun_out/{rnd}_launches.csv"):
        launches(f"gpurun_out/{rnd}_launches.csv", f"profiles/{rnd}_ncu_launch_shares.csv")
