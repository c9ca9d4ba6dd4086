"""Summarise the raw-metric CSV page of an `ncu --set full` report (ncu -i x.ncu-rep --page raw --csv): one block per profiled
launch with the metrics the roofline discussion uses (duration, DRAM bytes / throughput, tensor pipe, issue slots, registers)."""
import csv, json, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__cycles_elapsed.avg", "SM cycles")]


def main(path, traffic_out=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    traffic = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        print(f"== {name}  [{r[ix.get('ID', 0)]}]")
        vals = {}
        for k, label in KEYS:
            if k in ix:
                vals[k] = r[ix[k]]
                print(f"   {label:24s} {r[ix[k]]} {units[ix[k]]}")

        def num(k):
            try:
                v = float(r[ix[k]].replace(",", ""))
            except Exception:
                return None
            u = units[ix[k]].lower()
            mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return v * mul
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            traffic.append({"kernel": name, "grid": r[ix["launch__grid_size"]] if "launch__grid_size" in ix else None,
                            "time_under_ncu": r[ix["gpu__time_duration.sum"]] + " " + units[ix["gpu__time_duration.sum"]],
                            "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic": rd + wr})
    if traffic_out:
        json.dump(traffic, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
