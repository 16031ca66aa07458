"""Turn one round's GPU captures (scripts/gpu_profiles_r02.sh TAG) into the tracked evidence under profiles/:
python scripts/profile_summary.py TAG   (reads gpurun_out/TAG_{top.ncu-rep,launches.csv,bench.json})
 - profiles/TAG_top_kernels_ncu_summary.txt   key ncu metrics of the first launch of every kernel class
 - profiles/TAG_{fwd,bwd,chain}_source_hotspots.txt   per-CUDA-line stall tables
 - profiles/TAG_launches_fused_step.csv / _summary.txt, profiles/TAG_bench.json
 - profiles/traffic.json   DRAM bytes per launch (bench.py's roofline.traffic)"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    tag = sys.argv[1]
    g = lambda n: os.path.join(ROOT, "gpurun_out", n)
    p = lambda n: os.path.join(ROOT, "profiles", n)
    bench = json.loads(open(g(f"{tag}_bench.json")).read().strip().splitlines()[-1])
    E = float(bench.get("edges_start", bench["edges"]))
    out = subprocess.run(["ncu", "-i", g(f"{tag}_top.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["ncu --set full --clock-control none, bench.py --steps 2 --warmup 3 (cfg2: 269 beads x 128 molecules, W16A16).",
             f"First launch of every kernel class in the captured window; E = live directed edges (~{E / 1e6:.2f} M at this point of the run).",
             "Derived: warp-instructions per directed edge = smsp__inst_executed.sum / E.", ""]
    seen, traffic = {}, {}
    for r in rows[2:]:
        short = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        seen[short] = seen.get(short, 0) + 1
        if seen[short] > 1:
            continue
        lines.append(f"== {short}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k in WANT:
            if k in hdr:
                lines.append(f"   {k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        if "filter_cfconv" in short:
            lines.append(f"   {'warp-instructions per directed edge':72s} {float(r[hdr.index('smsp__inst_executed.sum')]) / E:16.1f}")
        i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic[short.split("<")[0]] = int(float(r[i]) * UNIT[units[i]] + float(r[j]) * UNIT[units[j]])
        lines.append("")
    open(p(f"{tag}_top_kernels_ncu_summary.txt"), "w").write("\n".join(lines))
    t = json.load(open(p("traffic.json")))
    for k in ("filter_cfconv_fwd_kernel", "filter_cfconv_bwd_kernel", "linear_chain_tc_kernel", "prior_csr_kernel",
              "edge_grad_to_forces_csr_kernel"):
        if k in traffic:
            t[k] = traffic[k]
    t["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures, cfg2 shapes. Fused-step kernels: "
                     f"profiles/{tag}_top_kernels_ncu_summary.txt; round-1 materialised / fp32-path kernels: profiles/r01a_cfconv_csr_ncu_summary.txt, "
                     "profiles/r01s_top_kernels_ncu_summary.txt, profiles/r01o_fp32_top_kernels_ncu_summary.txt")
    json.dump(t, open(p("traffic.json"), "w"), indent=1)
    for name, rx in (("fwd", "filter_cfconv_fwd"), ("bwd", "filter_cfconv_bwd"), ("chain", "linear_chain_tc")):
        o = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source.py"), g(f"{tag}_top.ncu-rep"), rx, "40"],
                           capture_output=True, text=True).stdout
        open(p(f"{tag}_{name}_source_hotspots.txt"), "w").write(o)
    shutil.copy(g(f"{tag}_launches.csv"), p(f"{tag}_launches_fused_step.csv"))
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"), g(f"{tag}_launches.csv"),
                    p(f"{tag}_launches_fused_step_summary.txt")], capture_output=True)
    shutil.copy(g(f"{tag}_bench.json"), p(f"{tag}_bench.json"))
    print("written:", sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith(tag)))


if __name__ == "__main__":
    main()
