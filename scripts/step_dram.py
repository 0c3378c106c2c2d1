"""Per-kernel duration and DRAM bytes of ONE step out of an ncu launch list.

    python scripts/step_dram.py profiles/r02_v_launches.csv > profiles/r02_v_step_dram.json

The launch list is the CSV log of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
around `bench.py` (cold cache, serialised launches: the SHARES are what bench.py compares with its live CUDA-event
family times, not the absolute durations).  One step = the launches after one `rank_topk_staged_kernel` up to and
including the next one; the last complete step of the log is taken.  `bench.py` copies the newest
`profiles/*_step_dram.json` into its roofline record (`hbm_view`).  Measurement infrastructure, not product code.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("made::", "").replace("<unnamed>::", "")
    return name.strip()


def main(path: str):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[hi]
    c_id, c_kn, c_mn, c_mu, c_mv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= c_mv:
            continue
        rec = launches.setdefault(r[c_id], dict(kernel=short(r[c_kn]), us=0.0, rd=0.0, wr=0.0))
        v = float(r[c_mv].replace(",", "")) * UNIT.get(r[c_mu], 1.0)
        if r[c_mn] == "gpu__time_duration.sum":
            rec["us"] = v
        elif r[c_mn] == "dram__bytes_read.sum":
            rec["rd"] = v
        elif r[c_mn] == "dram__bytes_write.sum":
            rec["wr"] = v
    seq = list(launches.values())
    marks = [i for i, l in enumerate(seq) if l["kernel"].startswith("rank_topk_staged_kernel")]
    if len(marks) < 2:
        raise SystemExit("need two rank_topk launches to delimit a step")
    step = seq[marks[-2] + 1:marks[-1] + 1]
    per = collections.OrderedDict()
    for l in step:
        p = per.setdefault(l["kernel"], dict(launches=0, us=0.0, dram_read_bytes=0.0, dram_write_bytes=0.0))
        p["launches"] += 1
        p["us"] += l["us"]
        p["dram_read_bytes"] += l["rd"]
        p["dram_write_bytes"] += l["wr"]
    fam = [k for k in per if k.startswith("gemm_tc_kernel") or k.startswith("ffn_fused_kernel")]
    g = dict(kernels=fam, launches=sum(per[k]["launches"] for k in fam), us=sum(per[k]["us"] for k in fam),
             dram_read_bytes=sum(per[k]["dram_read_bytes"] for k in fam),
             dram_write_bytes=sum(per[k]["dram_write_bytes"] for k in fam))
    g["dram_gbs_under_ncu"] = (g["dram_read_bytes"] + g["dram_write_bytes"]) / (g["us"] * 1e-6) / 1e9 if g["us"] else 0.0
    out = dict(source=f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                      "--clock-control none; one step = the launches between two rank_topk kernels; cold cache, serialised)",
               step_launches=len(step), step_kernel_us=sum(l["us"] for l in step),
               step_dram_bytes=sum(l["rd"] + l["wr"] for l in step), gemm_family=g)
    xp = [k for k in per if k.startswith("xpool_score_kernel")]
    if xp:
        out["xpool_score_kernel"] = per[xp[0]]
    out["per_kernel"] = dict(sorted(per.items(), key=lambda kv: -kv[1]["us"]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
