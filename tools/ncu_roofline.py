"""One roofline line per kernel of an ncu report: python tools/ncu_roofline.py rep.ncu-rep out.json
bound / achieved / peak / frac per kernel, with the bounding resource SURVEY 8(d) names for it, next to the ncu counters
the statement rests on (duration, DRAM bytes, FP64 pipe, issue slots, warps active)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]          # GB/s, measured copy bandwidth
FP64_WARP_INST = 148 * 4 * 0.5  # FP64 warp instructions per clock, whole chip (tools/cu/fp64_probe.cu: 0.5 / clk / scheduler)

# kernel-name prefix -> (bound named by SURVEY 8(d), what the kernel does)
BOUND = {
    "knn_cb_head_kernel": ("latency", "code-book search of the first locations of the ordering (window = everything before them)"),
    "knn_cb_kernel": ("issue", "code-book ordered search: one lane per location walks the (x+y)-sorted order (src/nn.cpp:136-211)"),
    "knn_brute_kernel": ("fp64", "brute-force ordered search: tiled pair distances (src/nn.cpp:63-111)"),
    "knn_grid_query_kernel": ("hbm", "query kNN over a uniform grid (RANN::nn2 stand-in)"),
    "tri_build_kernel": ("hbm", "resident distance triangle, written once per data set"),
    "q_resid_kernel": ("hbm", "Q(B, F, r, r) over stored factors (src/util.cpp:261-280)"),
    "q_forms_kernel": ("hbm", "all cross quadratic forms of [X y] in one pass"),
    "rep_solve_kernel": ("latency", "replicate recurrence, serial in the ordering (src/rNNGP.cpp:24-42): level-scheduled, one CTA"),
    "site_reduce_kernel": ("hbm", "per-site reductions of the prediction path"),
    "predict_emit_kernel": ("hbm", "posterior predictive draws: y0 = mean + sd z (src/rNNGPPredict.cpp:171-181)"),
    "bf_tri_kernel": ("fp64", "B/F + log-likelihood over the distance triangle"),
    "bf_ws_kernel": ("fp64", "B/F + log-likelihood, gather kernel"),
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        v = float(r[ci[k]].replace(",", ""))
        u = units[ci[k]]
        scale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        return v * scale.get(u, 1.0)

    res = []
    for r in rows[2:]:
        name = r[ci["Kernel Name"]]
        key = next((k for k in BOUND if k in name), None)
        dur = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        fp64 = float(r[ci["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]])
        issue = float(r[ci["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
        warps = float(r[ci["sm__warps_active.avg.pct_of_peak_sustained_active"]])
        d = {"kernel": name[:80], "duration_us": dur * 1e6, "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6,
             "fp64_pipe_active_pct": fp64, "issue_active_pct": issue, "warps_active_pct": warps,
             "grid": r[ci["launch__grid_size"]], "registers": r[ci["launch__registers_per_thread"]]}
        if key:
            bound, what = BOUND[key]
            d["what"] = what
            if bound == "hbm":
                ach = (rd + wr) / dur / 1e9
                d["roofline"] = {"bound": "hbm", "achieved": ach, "peak": HBM, "unit": "GB/s", "frac": ach / HBM,
                                 "note": "DRAM bytes of this launch / its duration, against the measured copy bandwidth"}
            elif bound == "fp64":
                d["roofline"] = {"bound": "fp64", "achieved": fp64 / 100 * FP64_WARP_INST, "peak": FP64_WARP_INST,
                                 "unit": "FP64 warp-instructions/clk", "frac": fp64 / 100,
                                 "note": "FP64 pipe busy fraction (ncu) = achieved / peak FP64 issue rate"}
            elif bound == "issue":
                d["roofline"] = {"bound": "issue slots", "achieved": issue, "peak": 100.0, "unit": "% of issue slots", "frac": issue / 100,
                                 "note": "integer / compare / shuffle bound walk: the scheduler's issue slots are the binding resource"}
            else:
                d["roofline"] = {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None,
                                 "note": "serial dependency chain: neither bandwidth nor a pipe is near its peak (see the counters)"}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        rf = d.get("roofline", {})
        print(f"{d['kernel'][:44]:44s} {d['duration_us']:10.1f} us  {rf.get('bound', '?'):12s} frac {rf.get('frac')}")


if __name__ == "__main__":
    main()
