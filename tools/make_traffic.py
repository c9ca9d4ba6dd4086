"""profiles/traffic_<tag>.json for bench.py's `roofline.traffic`: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one
representative launch per roofline class, from the raw-metric CSV pages of the two in-situ `ncu --set full` captures that
tools/gpu_evidence.sh takes (rollout and training bench runs).  usage: make_traffic.py <tag> <rollout.csv> <train.csv> [more train csv ...]"""
import csv, json, sys

# (workload, class) -> (kernel-name prefix, grid size or None, description, algorithmic bytes of that launch)
PICK = {
    ("rollout", "3"): ("block_tail_kernel<0, 2>", None, "fused block tail, inference, M=262144 (3 KB per token)", 262144 * 3072),
    ("rollout", "0"): ("gemm_tc_kernel<256, 1>", "147", "packed QKV projection, M=262144, N=768, K=256", 262144 * 2048),
    ("train", "3"): ("block_tail_kernel<1, 2>", None, "fused block tail, training stores, M=65536 (5.5 KB per token)", 65536 * 5632),
    ("train", "0"): ("gemm_tc_kernel<256, 1>", "148", "input-gradient GEMM, M=65536, N=K=256", 65536 * 1024),
    ("train", "2"): ("wgrad_tc_kernel<256>", "148", "weight gradient, M=65536, N=K=256", 65536 * 1024),
    ("train", "4"): ("mlp_bwd_kernel<2>", None, "fused MLP input-gradient chain, M=65536 (2 KB per token)", 65536 * 2048),
}


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    ix = {k: i for i, k in enumerate(hdr)}
    out = []
    for x in r[2:]:
        if len(x) != len(hdr):
            continue
        def num(k):
            v = float(x[ix[k]].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[ix[k]].lower(), 1)
        def us(k):
            v = float(x[ix[k]].replace(",", ""))
            return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[ix[k]].lower(), 1)
        out.append({"name": x[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("tante::", "").replace("(int)", "").strip(),
                    "grid": x[ix["launch__grid_size"]], "rd": num("dram__bytes_read.sum"), "wr": num("dram__bytes_write.sum"),
                    "us": us("gpu__time_duration.sum")})
    return out


def main(tag, roll_csv, *train_csvs):
    data = {"rollout": rows(roll_csv), "train": [r for f in train_csvs for r in rows(f)]}
    res = {"captured": f"round 2, tag {tag}", "source": "ncu --set full --clock-control none over `python bench.py` (rollout / training step), one B200; "
           "dram__bytes_read.sum + dram__bytes_write.sum of ONE representative launch per class (tools/gpu_evidence.sh, tools/make_traffic.py)"}
    for (wl, cls), (prefix, grid, desc, alg) in PICK.items():
        for r in data[wl]:
            if r["name"].startswith(prefix) and (grid is None or r["grid"].strip() == grid):
                res.setdefault(wl, {})[cls] = {"kernel": f"{r['name']}: {desc}", "time_us_under_ncu": round(r["us"], 1),
                                              "dram_read_bytes": int(r["rd"]), "dram_write_bytes": int(r["wr"]),
                                              "traffic": int(r["rd"] + r["wr"]), "algorithmic_bytes": alg}
                break
    json.dump(res, open(f"profiles/traffic_{tag}.json", "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
