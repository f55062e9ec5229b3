"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON + text file under profiles/.

    python tools/ncu_summary.py gpurun_out/full_fwd.ncu-rep profiles/r01_ncu_full_fwd.json [algorithmic_bytes] [flops]
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    alg_bytes = float(sys.argv[3]) if len(sys.argv) > 3 else None
    flops = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if u in UNIT_SCALE and ("bytes" in k or "duration" in k):
                    v *= UNIT_SCALE[u]
                    u = "byte" if "bytes" in k else "s"
                d[k] = {"value": v, "unit": u}
        t = d.get("gpu__time_duration.sum", {}).get("value")
        rd = d.get("dram__bytes_read.sum", {}).get("value", 0.0)
        wr = d.get("dram__bytes_write.sum", {}).get("value", 0.0)
        d["derived"] = {"dram_bytes": rd + wr, "dram_GBps": (rd + wr) / t / 1e9 if t else None,
                        "algorithmic_bytes": alg_bytes, "traffic_over_algorithmic": (rd + wr) / alg_bytes if alg_bytes else None,
                        "tflops_under_ncu": flops / t / 1e12 if (flops and t) else None}
        kernels.append(d)
    json.dump({"report": rep, "note": "captured with ncu --set full --clock-control none; times under the profiler are "
               "not bench values", "kernels": kernels}, open(out, "w"), indent=1)
    for d in kernels:
        print(d["kernel"][:120])
        for k, v in d.items():
            if isinstance(v, dict) and "value" in v:
                print(f"   {k:75s} {v['value']:.6g} {v['unit']}")
        print("   derived", d["derived"])


if __name__ == "__main__":
    main()
