"""Turn the ncu reports of a GPU session (gpurun_out/<tag>_*) into the tracked text summaries under
profiles/ (needs the `ncu` CLI; run in the build container after tools/gpu_session.sh).

    python tools/make_profiles.py <tag> <round-name>
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEYS  # noqa: E402

EXTRA = ["dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_active.avg",
         "sm__cycles_active.max", "sm__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle_active",
         "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
         "launch__shared_mem_per_block_static", "sm__warps_active.avg.per_cycle_active",
         "smsp__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
         "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_st.sum",
         "sm__sass_inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_shared_ld.sum",
         # unit utilisations: which pipe is the busiest
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
         "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def sass_blocks(rep, top=12):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_blocks.py"), rep, "1", str(top)],
                         capture_output=True, text=True).stdout
    return out


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    traffic = {}
    for kern in ("advance_kernel", "render_tile_kernel"):
        rep = os.path.join(src, "%s_full_%s.ncu-rep" % (tag, kern))
        if not os.path.exists(rep):
            continue
        rows, units = raw_page(rep)
        with open(os.path.join(dst, "%s_ncu_%s.txt" % (rnd, kern)), "w") as fh:
            fh.write("# ncu --set full --import-source on --clock-control none -k regex:%s -c 1 (after the warm-up launches)\n" % kern)
            fh.write("# python tools/perf_probe.py --worlds 131072 --balls 8 --steps 12 --render --warm 8 --max-substeps 64\n")
            fh.write("# (one steady-state launch; times under ncu are serialised and cold-cache)\n")
            for d in rows:
                fh.write("== %s grid %s block %s\n" % (d["Kernel Name"], d.get("Grid Size"), d.get("Block Size")))
                for k in KEYS + EXTRA:
                    if k in d:
                        fh.write("   %-88s %s %s\n" % (k, d[k], units.get(k, "")))
                rd = float(d["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_write.sum"]]
                traffic[kern] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                                 "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"]) * (1e3 if units["gpu__time_duration.sum"] == "ms" else 1.0),
                                 "source": "profiles/%s_ncu_%s.txt" % (rnd, kern)}
            fh.write("\n# heaviest SASS runs (instructions executed per launch)\n")
            fh.write(sass_blocks(rep))
    for name in ("launches.csv", "bench.json", "bench_reference.json", "pytest_gpu.log", "smoke.log", "sweep.json", "store_ceiling.log",
                 "steps.log", "native.log", "e2e.log"):
        p = os.path.join(src, "%s_%s" % (tag, name))
        if os.path.exists(p):
            shutil.copy(p, os.path.join(dst, "%s_%s" % (rnd, name)))
    with open(os.path.join(dst, "traffic.json"), "w") as fh:
        json.dump({"round": rnd, "workload": "131072 worlds x 8 balls, 64x64 frames, fp32", "kernels": traffic}, fh, indent=1)
    # share of each kernel in the launch list
    p = os.path.join(src, "%s_launches.csv" % tag)
    if os.path.exists(p):
        tot = {}
        for r in csv.DictReader(l for l in open(p) if l.startswith('"')):
            if r.get("Metric Name") == "gpu__time_duration.sum":
                k = r["Kernel Name"].split("(")[0].replace("void ", "")
                a = tot.setdefault(k, [0, 0.0])
                a[0] += 1
                a[1] += float(r["Metric Value"])
        s = sum(v[1] for v in tot.values())
        with open(os.path.join(dst, "%s_launch_shares.txt" % rnd), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 40: python bench.py --steps 6 --warmup 3 (setup, warm-up, the timed steps, the second pass with the advance events, the render-alone loop)\n")
            for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                fh.write("%-40s launches %3d  total %10.1f us  avg %9.1f us  share %5.1f%%\n" % (k, n, t / 1e3, t / 1e3 / n, 100 * t / s))
        print(open(os.path.join(dst, "%s_launch_shares.txt" % rnd)).read())
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
