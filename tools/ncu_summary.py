"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.

  python tools/ncu_summary.py prof.ncu-rep [--json profiles/rN_ncu_traffic.json] [--mode train|render]

--json writes {kernel class: mean dram__bytes_read.sum + dram__bytes_write.sum per launch} -- what bench.py reports as
`roofline.traffic` (it reads the newest profiles/r*_ncu_traffic.json)."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def kernel_class(name, mode):
    if "nerf_mlp_fwd_tc_kernel<1" in name or "nerf_mlp_fwd_tc_kernel<true" in name:
        return "mlp_fwd_train"
    if "nerf_mlp_fwd_tc_kernel" in name:
        return "mlp_fwd_render"
    if "nerf_mlp_bwd_tc_kernel" in name:
        return "mlp_bwd_chain"
    if "nerf_wgrad_tc_kernel" in name:
        return "wgrad"
    if "nerf_input_grad_tc_kernel" in name:
        return "input_grad"
    return None


def main(argv):
    path = argv[0]
    js = argv[argv.index("--json") + 1] if "--json" in argv else None
    mode = argv[argv.index("--mode") + 1] if "--mode" in argv else "train"
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    traffic = {}
    for r in data:
        print("kernel:", r[name_i][:100])
        tot = 0.0
        for i, h in enumerate(hdr):
            short = h.split("TriageCompute.")[-1]
            if short in KEYS:
                print(f"  {short:80s} {r[i]:>16s} {units[i]}")
            if short in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                try:
                    tot += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    pass
        print()
        k = kernel_class(r[name_i], mode)
        if k:
            traffic.setdefault(k, []).append(tot)
    if js:
        prev = {}
        try:
            prev = json.load(open(js))
        except Exception:
            pass
        prev.update({k: sum(v) / len(v) for k, v in traffic.items()})
        prev["_unit"] = "bytes per launch, mean of the coarse (64 samples/ray) and fine (192) launches of a 4096-ray step"
        prev.setdefault("_source", [])
        if path not in prev["_source"]:
            prev["_source"].append(path)
        json.dump(prev, open(js, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
