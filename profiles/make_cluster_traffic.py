#!/usr/bin/env python
"""profiles/cluster_traffic.json (read by bench.py for roofline.traffic) from the ncu traffic pass of
profiles/run_ncu.sh:  python profiles/make_cluster_traffic.py profiles/<tag>_traffic.csv <tag>
(the L1TEX / issue figures of profiles/<tag>_cluster_ncu_raw.csv are added as "onchip_pipe" when that file exists)"""
import csv, json, os, re, sys
path, tag = sys.argv[1], sys.argv[2]
per = {}
for r in csv.reader(open(path)):
    if len(r) > 14 and r[0].isdigit():
        k = re.sub(r"^void\s+|\(.*", "", r[4])
        per.setdefault(k, {})[r[12]] = float(r[14])
rd = sum(v["dram__bytes_read.sum"] for v in per.values())
wr = sum(v["dram__bytes_write.sum"] for v in per.values())
out = {
    "kernel": "k_pcg_cluster<1..4> (the four concurrent persistent kernels of one device-resident step, profiled one after the other)",
    "source": "profiles/%s_traffic.csv (ncu --metrics dram__bytes_*.sum,lts__t_sectors.sum,gpu__time_duration.sum) and "
              "profiles/%s_cluster_ncu_raw.csv (--set full)" % (tag, tag),
    "workload": "100 plates x 4 conditions, mesh_size 1e-2", "nnz": 49452160, "n_active_dofs": 3664196,
    "per_kernel": per, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
    "dram_bytes_per_launch": rd + wr,
    "l2_bytes_per_launch": 32.0 * sum(v["lts__t_sectors.sum"] for v in per.values()),
}
raw = os.path.join(os.path.dirname(os.path.abspath(__file__)), "%s_cluster_ncu_raw.csv" % tag)
if os.path.exists(raw):
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    kcol = hdr.index("Kernel Name") if "Kernel Name" in hdr else 0
    pipe = {}
    for r in rows[2:]:
        pipe[re.sub(r"^void\s+|\(.*", "", r[kcol])] = {
            "l1tex_pct_of_peak_elapsed": float(r[hdr.index("l1tex__throughput.avg.pct_of_peak_sustained_elapsed")]),
            "sm_active_frac": float(r[hdr.index("sm__cycles_active.avg")]) / float(r[hdr.index("sm__cycles_elapsed.avg")]),
            "issue_active_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
            "shared_wavefronts_pct_of_peak_elapsed": float(r[hdr.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")]),
            "shared_ld_bank_conflict_wavefronts": float(r[hdr.index("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum")]),
            "shared_wavefronts": float(r[hdr.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")])}
    out["onchip_pipe"] = {"per_kernel": pipe, "note": "L1TEX data pipe: pct of peak over the launch; divide by sm_active_frac for the "
                          "share while the SMs hold a CTA (DESIGN 3.1a)"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cluster_traffic.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_launch", "l2_bytes_per_launch")}))
