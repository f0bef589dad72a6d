"""Builds profiles/r2_stage_table_<G>k.json (embedded by bench.py as roofline.stages) and the per-size entry of
profiles/roofline_traffic.json from ONE ncu pass over an eager steady-state iteration:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none --csv --log-file gpurun_out/stages_100k.csv python tools/prof_iteration.py 100000 3

  python tools/stage_table.py gpurun_out/stages_100k.csv 100000 [hbm_gbs]

The LAST complete iteration of the list is used (from one gsd_track_node_prep launch to the next).  Per launch: device time
(cold-cache, serialised: compare SHARES), DRAM bytes, warp instructions, and the two fractions that matter: DRAM bytes / time
against the measured HBM peak, warp instructions / time against the issue-slot peak (148 SMs x 4 schedulers x SM clock)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, G = sys.argv[1], int(sys.argv[2])
    hbm = float(sys.argv[3]) if len(sys.argv) > 3 else float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(r for r in rows if "Kernel Name" in r)
    st = rows.index(h) + 1
    kid, kn, mn, mv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    grid, block = h.index("Grid Size"), h.index("Block Size")
    launches = {}
    order = []
    for r in rows[st:]:
        if len(r) <= mv:
            continue
        i = int(r[kid])
        if i not in launches:
            launches[i] = {"kernel": re.sub(r"\(.*", "", r[kn]).replace("void ", "")[:80], "grid": r[grid], "block": r[block]}
            order.append(i)
        try:
            launches[i][r[mn]] = float(r[mv].replace(",", ""))
        except ValueError:
            pass
    L = [launches[i] for i in order]
    # first launch of an eager iteration: the priors' node-record kernel (side branch, launched first by the host); older builds
    # started with the rotation normalisation kernel
    idx = [i for i, l in enumerate(L) if "track_node_prep" in l["kernel"]] or [i for i, l in enumerate(L) if "normalize_rot" in l["kernel"]]
    if len(idx) < 2:
        raise SystemExit("need at least two iterations in the launch list")
    it = L[idx[-2]:idx[-1]]
    sm_mhz = 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6
    stages, tot = [], 0.0
    for l in it:
        t = l.get("gpu__time_duration.sum", 0.0) * 1e-9
        d = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        w = l.get("smsp__inst_executed.sum", 0.0)
        tot += t
        stages.append({"kernel": l["kernel"], "grid": l["grid"], "block": l["block"], "us": round(t * 1e6, 2), "dram_bytes": d,
                       "hbm_frac": round(d / t / 1e9 / hbm, 4) if t else None, "warp_insts": w,
                       "issue_frac": round(w / t / issue_peak, 4) if t else None})
    for s in stages:
        s["share"] = round(s["us"] * 1e-6 / tot, 4)
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none "
                     "over tools/prof_iteration.py %d (eager steady-state iteration; per-launch times are cold-cache and serialised)" % G,
           "gaussians": G, "hbm_peak_gbs": hbm, "issue_peak_warp_inst_per_s": issue_peak, "sm_mhz_assumed": sm_mhz,
           "serialised_sum_us": round(tot * 1e6, 1), "launches": len(stages), "stages": stages}
    dst = os.path.join(ROOT, "profiles", "r2_stage_table_%dk.json" % (G // 1000))
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)
    for s in stages:
        print("%7.1f us %5.1f%%  hbm %5.1f%%  issue %5.1f%%  %s" % (s["us"], 100 * s["share"], 100 * (s["hbm_frac"] or 0), 100 * (s["issue_frac"] or 0), s["kernel"]))
    print("%7.1f us total, %d launches" % (tot * 1e6, len(stages)))
    # the dominant kernel's entry for bench.py's roofline.traffic / issue_slot
    dom = [s for s in stages if "blend_bwd_chunk" in s["kernel"]]
    if dom:
        rt = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        cur = json.load(open(rt)) if os.path.exists(rt) else {}
        cur[str(G)] = {"blend_backward_dram_bytes_per_launch": dom[0]["dram_bytes"], "blend_backward_warp_instructions_per_launch": dom[0]["warp_insts"],
                       "source": "profiles/r2_stage_table_%dk.json" % (G // 1000)}
        json.dump(cur, open(rt, "w"), indent=1)
        print("updated", rt)


if __name__ == "__main__":
    main()
