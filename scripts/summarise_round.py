"""After `bash scripts/gpu_round.sh <tag>` on the GPU box: turn gpurun_out/<tag>_* into the tracked summaries under profiles/
and refresh profiles/roofline_inputs.json (the ncu-derived figures bench.py reports).   python scripts/summarise_round.py <tag>"""
import csv, json, os, subprocess, sys
tag = sys.argv[1]
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
rep = os.path.join(G, tag + "_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
open("/tmp/%s_raw.csv" % tag, "w").write(raw); open("/tmp/%s_src.csv" % tag, "w").write(src)
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(n):
    i = hdr.index(n); return [float(d[i].replace(",", "")) for d in data], units[i]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
dr, u1 = col("dram__bytes_read.sum"); dw, u2 = col("dram__bytes_write.sum"); inst, _ = col("smsp__inst_executed.sum")
lanes, _ = col("smsp__thread_inst_executed_per_inst_executed.ratio"); iss, _ = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
bench_line = [l for l in open(os.path.join(G, tag + "_ncu_f.log")) if l.startswith('{"metric')][-1]
exp = json.loads(bench_line)["plan"]["expansions_per_plan"]
name = "%s_expand_kernel_ncu_full.txt" % tag
out = {"_comment": "Per-launch figures of kgmt::expand_kernel<grid/smem, no-record> on BASELINE config 2, read from the ncu --set full capture summarised in profiles/%s (mean of the captured launches). bench.py copies them into roofline.traffic and roofline_issue." % name,
       "source": "profiles/" + name, "workload": "c2", "collision": "grid",
       "dram_bytes_per_launch": int(sum(a * scale[u1] + b * scale[u2] for a, b in zip(dr, dw)) / len(dr)),
       "warp_instructions_per_launch": int(sum(inst) / len(inst)), "expansions_per_launch": int(exp),
       "warp_instructions_per_expansion": round(sum(inst) / len(inst) / exp, 1),
       "issue_active_pct": round(sum(iss) / len(iss), 1), "avg_active_lanes": round(sum(lanes) / len(lanes), 1)}
json.dump(out, open(os.path.join(P, "roofline_inputs.json"), "w"), indent=1)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "lts__t_bytes.sum"]
with open(os.path.join(P, name), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 2, python bench.py --steps 2 --warmup 1 (config C2), pass %s\n" % tag)
    f.write("# kgmt::expand_kernel<grid/smem, no-record>; captured launches in columns\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w); f.write("%-75s %-16s %s\n" % (w, units[i], "  ".join(d[i] for d in data)))
    f.write("\n# source page (first launch): hottest source lines — share of warp instructions (I), average active lanes, share of stall samples (S), top stall reasons\n")
    f.write(subprocess.run([sys.executable, os.path.join(R, "scripts", "ncu_hot_lines.py"), "/tmp/%s_src.csv" % tag, "22"], capture_output=True, text=True).stdout)
    f.write("\n# derived (profiles/roofline_inputs.json): " + json.dumps({k: v for k, v in out.items() if not k.startswith("_")}) + "\n")
subprocess.run(["cp", os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches_bench_c2.csv")])
with open(os.path.join(P, tag + "_bench_baselines_timeline.txt"), "w") as f:
    def tail(fn, n=1):
        try:
            return "".join(open(os.path.join(G, fn)).readlines()[-n:])
        except OSError:
            return "(missing)\n"
    f.write("# gpurun pass %s, one B200; host CPU: %s" % (tag, tail(tag + "_smi.log", 2).replace("\n", " ") + "\n"))
    f.write("## bench.py (default: C2, 20 steps of 64 plans, 3 warm-up)\n" + tail(tag + "_bench.log"))
    f.write("## bench.py --impl reference --steps 3 --warmup 1\n" + tail(tag + "_bench_ref.log"))
    f.write("## baseline A: the reference's own CUDA kernels recompiled for sm_100a (scripts/baseline_ref_gpu.py)\n")
    if os.path.exists(os.path.join(G, tag + "_baselineA.log")):
        f.write("".join(l for l in open(os.path.join(G, tag + "_baselineA.log")) if l.startswith(("c1 P", "c2 P", "reference plan", "ours plan"))))
    else:
        f.write("(in the bench line: same_population / ref_cuda_baseline / ttfs.reference_ms)\n")
    f.write("## per-iteration device timeline of one C2 plan (scripts/iter_profile.py timeline)\n" + open(os.path.join(G, tag + "_timeline.log")).read())
    f.write("## pytest -m gpu\n" + tail(tag + "_pytest.log", 2))
print(json.dumps(out, indent=1))
