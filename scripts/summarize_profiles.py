#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py gpurun_out/prof_ray_r1_final.ncu-rep profiles/r1_final_launches.csv

Reads the .ncu-rep with `ncu -i ... --page raw|source --csv` (no GPU needed) and writes
profiles/r1_final_ray_ncu.md, profiles/r1_final_launches.md and profiles/ncu_traffic.json.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_lookup_hit.sum', 'lts__t_sectors_lookup_miss.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v.replace(",", "")) * mult


def ray_summary(rep):
    raw = page(rep, "raw")
    hdr, units, r = raw[0], raw[1], raw[2]
    vals = {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr}
    src = [x for x in page(rep, "source")[2:] if len(x) >= 10 and x[0] not in ("Kernel Name", "Address")]
    tot = sum(int(x[5]) for x in src)
    rays = 1048576 * 60 / 32.0
    b = collections.OrderedDict([("march loop, skip path", 0), ("march loop, near-wall path", 0),
                                 ("per ray (direction, table product)", 0), ("per particle (sincos, pow, window staging)", 0)])
    for x in src:
        ie = int(x[5])
        if ie > 10 * rays:
            b["march loop, skip path"] += ie
        elif ie > 4 * rays:
            b["march loop, near-wall path"] += ie
        elif ie > 0.6 * rays:
            b["per ray (direction, table product)"] += ie
        else:
            b["per particle (sincos, pow, window staging)"] += ie
    dram = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
    sass = " ".join(x[1] for x in src)
    md = ["# k_raycast_weight<8, 207> -- ncu --set full, final round-1 kernel", "",
          "Command (under gpurun, after the same command exited 0 without ncu):", "",
          "    ncu --set full --clock-control none --import-source on -k regex:k_raycast_weight -s 3 -c 1 \\",
          "        -o gpurun_out/prof_ray_r1_final python bench.py --steps 3 --warmup 3 --no-cpu", "",
          "Workload: Spielberg_map, 1,048,576 particles x 60 beams (62.9 M rays), tracking cloud, 4th update of the run.",
          "Times under ncu are cold-cache and serialised: compare shares, not absolutes.", "",
          "| metric | value | unit |", "|---|---|---|"]
    md += ["| `%s` | %s | %s |" % (k, v[0], v[1]) for k, v in vals.items()]
    md += ["", "DRAM traffic per launch: %.1f MB (particle state in, weights out, the window staged once per CTA mostly from L2)." % (dram / 1e6),
           "Algorithmic bytes per launch (SURVEY 8d: N*R*C-bar + 32 N): ~5.5 GB at C-bar 86 -- the bytes the reference's march reads from",
           "its int8 grid; here they are replaced by shared-memory lookups of the skip map (%s shared-memory wavefronts, %s of them bank-conflict replays)." % (
               vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0], vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0]),
           "", "SASS evidence: UBLKCP (cp.async.bulk window staging) %s, SYNCS (mbarrier) %s, LDS.U8 lookups %s; no HMMA/UTC*MMA (no dense contraction on this path)." % (
               "present" if "UBLKCP" in sass else "absent", "present" if "SYNCS" in sass else "absent", "present" if "LDS.U8" in sass else "absent"),
           "", "Executed warp instructions by region (source page, `Instructions Executed`):", "",
           "| region | warp instructions | share | per warp-ray |", "|---|---|---|---|"]
    md += ["| %s | %.3e | %.1f %% | %.0f |" % (k, v, 100 * v / tot, v / rays) for k, v in b.items()]
    md += ["| total | %.3e | | %.0f |" % (tot, tot / rays), "",
           "Reading: warp-issue bound (issue active %s %% of peak; tensor, FP64 and LSU pipes nearly idle), SIMT lane efficiency %s of 32 after the"
           % (vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0][:4], vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
           "heading sort; the march loop is ~90 % of all instructions."]
    open(os.path.join(ROOT, "profiles", "r1_final_ray_ncu.md"), "w").write("\n".join(md) + "\n")
    json.dump({"k_raycast_weight_dram_bytes_per_launch": dram,
               "source": "profiles/r1_final_ray_ncu.md (ncu --set full, one launch, 1M x 60 Spielberg)"},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"))


def launch_summary(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[kn].split("(")[0][:60], []).append(v)
    upd = {k: v for k, v in agg.items() if "mclb200" in k and not any(x in k for x in ("k_range_queries", "k_init_pose", "k_fill"))}
    ray_key = [k for k in upd if "k_raycast_weight" in k][0]
    n_upd = len(upd[ray_key])
    tot = sum(sum(v) for v in upd.values())
    out = ["# Launch list of `python bench.py --steps 3 --warmup 3 --no-cpu`", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`; raw CSV: `profiles/r1_final_launches.csv`.",
           "Per-launch times are cold-cache and serialised by ncu; the shares are what must (and do) agree with the CUDA-event",
           "stage times in the bench line (ray+weight stage 0.818 of 1.031 ms = 79 %).", "",
           "Kernels of one MCL update (%d updates captured):" % n_upd, "",
           "| kernel | launches / update | mean us / launch | share of update |", "|---|---|---|---|"]
    for k, v in upd.items():
        out.append("| `%s` | %.0f | %.1f | %.1f %% |" % (k.replace("void ", "").replace("mclb200::", ""), len(v) / n_upd,
                                                        sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    out.append("| total per update | %d | %.1f | |" % (round(sum(len(v) for v in upd.values()) / n_upd), tot / n_upd / 1e3))
    rq = agg.get("mclb200::k_range_queries", [0])
    out += ["", "Outside the update: `k_range_queries` (synthetic scan generation, %d launches, %.1f us each), `k_init_pose`, `k_fill`, "
            "and the L2-flush fill kernel of bench.py." % (len(rq), sum(rq) / max(1, len(rq)) / 1e3)]
    open(os.path.join(ROOT, "profiles", "r1_final_launches.md"), "w").write("\n".join(out) + "\n")


def dir_summary(rep, launches_csv, tag):
    """Summaries for the directional ray stage (k_raycast_dir): profiles/<tag>_ray_ncu.md,
    profiles/<tag>_launches.md (+ .csv copy) and the kernel's DRAM traffic in ncu_traffic.json."""
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    r = [x for x in raw[2:] if "k_raycast_dir" in x[hdr.index("Kernel Name")]][0]
    vals = {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr}
    kname = r[hdr.index("Kernel Name")]
    src_rows = page(rep, "source")
    # the source page lists every captured kernel; take the first k_raycast_dir block
    src, take = [], False
    for x in src_rows:
        if x and x[0] == "Kernel Name":
            if take:
                break
            take = "k_raycast_dir" in x[1]
            continue
        if take and len(x) >= 10 and x[0] != "Address":
            src.append(x)
    tot = sum(int(x[5]) for x in src)
    rays = 1048576 * 60 / 32.0
    b = collections.OrderedDict([("march loop (branch-free body, one trip per lookup)", 0),
                                 ("per warp-ray (beam direction, sector test, result store, loop entry/exit)", 0),
                                 ("per warp-unit (record load + prefetch, window base, first-sample lookup, beam range)", 0),
                                 ("rare (window staging, unit counter, exact replay)", 0)])
    samples = collections.OrderedDict((k, 0) for k in b)
    lanes_loop = [0, 0]
    for x in src:
        ie = int(x[5])
        q = ie / rays
        key = (list(b)[0] if q >= 3 else list(b)[1] if q >= 0.7 else list(b)[2] if q >= 0.2 else list(b)[3])
        b[key] += ie
        samples[key] += int(x[4])
        if q >= 3:
            lanes_loop[0] += int(x[6])
            lanes_loop[1] += ie
    nsamp = max(1, sum(samples.values()))
    dram = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
    sass = " ".join(x[1] for x in src)
    md = ["# %s -- ncu --set full (directional ray stage)" % kname, "",
          "Command (under gpurun, after the same command exited 0 without ncu):", "",
          "    ncu --set full --clock-control none --import-source on -k regex:k_raycast_dir -s 6 -c 1 -f \\",
          "        -o gpurun_out/prof_%s python bench.py --steps 3 --warmup 3 --no-cpu" % tag, "",
          "Workload: Spielberg_map, 1,048,576 particles x 60 beams (62.9 M rays), tracking cloud.",
          "Times under ncu are cold-cache and serialised: compare shares, not absolutes.", "",
          "| metric | value | unit |", "|---|---|---|"]
    md += ["| `%s` | %s | %s |" % (k, v[0], v[1]) for k, v in vals.items()]
    md += ["", "DRAM traffic per launch: %.1f MB (ray-start records in, step bytes out, sector windows mostly from L2)." % (dram / 1e6),
           "Algorithmic bytes per launch (SURVEY 8d): N*R*C-bar grid bytes of the reference's march + 64 B/particle records + 1 B/ray steps.",
           "Shared-memory lookups of the sector windows: %s wavefronts, %s of them bank-conflict replays." % (
               vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0], vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0]),
           "", "SASS evidence: UBLKCP (cp.async.bulk window staging) %s, SYNCS (mbarrier) %s, LDGSTS (cp.async record prefetch) %s, LDS.U8 lookups %s; "
           "no HMMA/UTC*MMA (no dense contraction on this path)." % tuple(
               "present" if t in sass else "absent" for t in ("UBLKCP", "SYNCS", "LDGSTS", "LDS.U8")),
           "", "Executed warp instructions by region (source page, `Instructions Executed`):", "",
           "| region | warp instructions | share | per warp-ray | share of the stall samples |", "|---|---|---|---|---|"]
    md += ["| %s | %.3e | %.1f %% | %.0f | %.1f %% |" % (k, v, 100 * v / tot, v / rays, 100.0 * samples[k] / nsamp) for k, v in b.items()]
    md += ["| total | %.3e | | %.0f | |" % (tot, tot / rays), "",
           "Active lanes per instruction inside the march loop: %.1f of 32 (whole kernel: %s)." % (
               lanes_loop[0] / max(1, lanes_loop[1]), vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
           "", "Reading: warp-issue bound (issue active %s %% of peak over the whole launch, i.e. including window staging, barrier waits and the tail; "
           "tensor, FP64 and LSU pipes far from saturated)." % vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0][:4]]
    stalls = [(k.split("issue_stalled_")[1].split("_per_issue")[0], float(r[hdr.index(k)])) for k in hdr
              if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    stalls.sort(key=lambda kv: -kv[1])
    md += ["", "Warp states per issued instruction (ncu `smsp__average_warps_issue_stalled_*_per_issue_active`): " +
           ", ".join("%s %.2f" % kv for kv in stalls[:8]) + "."]
    open(os.path.join(ROOT, "profiles", tag + "_ray_ncu.md"), "w").write("\n".join(md) + "\n")
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    t = json.load(open(tp)) if os.path.exists(tp) else {}
    t["k_raycast_dir_dram_bytes_per_launch"] = dram
    t["k_raycast_dir_warp_instructions_per_launch"] = float(vals["smsp__inst_executed.sum"][0].replace(",", ""))
    t["k_raycast_dir_source"] = "profiles/%s_ray_ncu.md (ncu --set full, one launch, 1M x 60 Spielberg)" % tag
    t["k_raycast_dir_threads_per_instruction"] = float(vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
    wf = float(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0].replace(",", ""))
    t["k_raycast_dir_shared_conflict_share"] = float(vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0].replace(",", "")) / max(1.0, wf)
    # the configuration the capture belongs to: bench.py only quotes it for a run of the same workload
    t.update({"map": "Spielberg_map", "particles": 1048576, "beams": 60, "capture": tag})
    json.dump(t, open(tp, "w"))
    # launch list
    import shutil
    shutil.copy(launches_csv, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
    rows = list(csv.reader(open(launches_csv)))
    hi = [i for i, rr in enumerate(rows) if "Kernel Name" in rr][0]
    h2 = rows[hi]
    kn, mv = h2.index("Kernel Name"), h2.index("Metric Value")
    agg = collections.OrderedDict()
    for rr in rows[hi + 2:]:
        if len(rr) <= mv:
            continue
        try:
            v = float(rr[mv].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(rr[kn].split("(")[0][:60], []).append(v)
    skip = ("k_range_queries", "k_init_pose", "k_fill", "k_build_dir_maps", "k_gather_bench", "k_map_masks", "k_edt_cols", "k_edt_rows", "k_map_codes")
    upd = {k: v for k, v in agg.items() if "mclb200" in k and not any(x in k for x in skip)}
    n_upd = len(upd[[k for k in upd if "k_raycast_dir" in k][0]])
    # the diagnostics pass of bench.py (per-ray steps kept) inflates k_weight_steps: use the median
    med = {k: sorted(v)[len(v) // 2] for k, v in upd.items()}
    tot_med = sum(med[k] * len(v) / n_upd for k, v in upd.items())
    out = ["# Launch list of `python bench.py --steps 3 --warmup 3 --no-cpu` (directional ray stage)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv`; raw CSV: `profiles/%s_launches.csv`." % tag,
           "Per-launch times are cold-cache and serialised by ncu; the SHARES are what must agree with the CUDA-event stage",
           "times of the bench line.  Median per launch (bench.py's C-bar pass keeps per-ray steps in a few updates, which",
           "inflates `k_weight_steps` there).", "",
           "Kernels of one MCL update (%d updates captured):" % n_upd, "",
           "| kernel | launches / update | median us / launch | share of update |", "|---|---|---|---|"]
    for k, v in upd.items():
        out.append("| `%s` | %.0f | %.1f | %.1f %% |" % (k.replace("void ", "").replace("mclb200::", ""), len(v) / n_upd,
                                                        med[k] / 1e3, 100 * med[k] * len(v) / n_upd / tot_med))
    out.append("| total per update | %d | %.1f | |" % (round(sum(len(v) for v in upd.values()) / n_upd), tot_med / 1e3))
    bd = agg.get("mclb200::k_build_dir_maps", [0])
    edt = sum(agg.get("mclb200::" + k, [0])[0] for k in ("k_map_masks", "k_edt_cols", "k_edt_rows", "k_map_codes"))
    out += ["", "Outside the update: `k_build_dir_maps` once per map (%.1f ms for the 16 sector maps of Spielberg_map), the isotropic skip map "
            "(`k_map_masks`, `k_edt_cols`, `k_edt_rows`, `k_map_codes`: %.1f ms, exact Euclidean transform on the device), `k_range_queries` "
            "(synthetic scan generation), `k_init_pose`, `k_fill`, the gather micro-benchmark and the L2-flush fill of bench.py." % (bd[0] / 1e6, edt / 1e6)]
    open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w").write("\n".join(out) + "\n")
    print("wrote profiles/%s_ray_ncu.md, profiles/%s_launches.md, profiles/ncu_traffic.json" % (tag, tag))


if __name__ == "__main__" and not (len(sys.argv) >= 3 and sys.argv[1] == "--small"):
    if len(sys.argv) >= 4 and sys.argv[1] == "--dir":
        dir_summary(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "r1_dir")
    else:
        ray_summary(sys.argv[1])
        launch_summary(sys.argv[2])
        print("wrote profiles/r1_final_ray_ncu.md, profiles/r1_final_launches.md, profiles/ncu_traffic.json")


def small_summary(rep, tag):
    """profiles/<tag>_small_ncu.md: one row per captured launch of the streaming kernels (ncu --set full)."""
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    cols = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes")]
    cols = [(k, n) for k, n in cols if k in hdr]
    kn = hdr.index("Kernel Name")
    seen, rows = {}, []
    for r in raw[2:]:
        name = r[kn].replace("void ", "").replace("mclb200::", "").split("(")[0]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 2:
            continue
        cells = []
        for k, _ in cols:
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            if "byte" in u:
                cells.append("%.1f MB" % (to_bytes(v, u) / 1e6))
            else:
                try:
                    cells.append("%.1f" % float(v.replace(",", "")) if "." in v else v)
                except ValueError:
                    cells.append(v)
        rows.append("| `%s` | %s |" % (name, " | ".join(cells)))
    md = ["# Streaming kernels of the update -- ncu --set full", "",
          "    ncu --set full --clock-control none --import-source on -k regex:'k_resample_motion|k_weight_steps|k_exact_pass|k_exact_emit|k_tile_sums|k_dir_gather|k_sort' \\",
          "        -s 40 -c 12 -f -o gpurun_out/prof_%s_small python bench.py --steps 3 --warmup 3 --no-cpu" % tag, "",
          "Workload: Spielberg_map, 1,048,576 particles x 60 beams.  Cold-cache, serialised launches (ncu): the CUDA-event times of the",
          "bench line (`kernels`) are the timings; this table says what each kernel is bound by.", "",
          "| kernel | " + " | ".join(n for _, n in cols) + " |", "|---|" + "---|" * len(cols)] + rows
    open(os.path.join(ROOT, "profiles", tag + "_small_ncu.md"), "w").write("\n".join(md) + "\n")
    print("wrote profiles/%s_small_ncu.md" % tag)


if __name__ == "__main__" and len(sys.argv) >= 3 and sys.argv[1] == "--small":
    small_summary(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "r2")
