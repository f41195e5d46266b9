"""Summarise an `ncu --set full` report (run here, where ncu can read .ncu-rep files) into the small JSON files kept
under profiles/: per-kernel duration, DRAM bytes, pipe utilisation, issue activity, L2 hit rate.
usage: python tools/ncu_extract.py profiles/r01_ncu_final_kernels.json profiles/r01_traffic.json gpurun_out/final.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "tc_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_cycles_pct",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_issue_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "smem_lsu_wavefronts_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tc_wavefronts_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__inst_executed.sum": "warp_instructions",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed": "gcc_instruction_requests_pct",
    "sm__icc_request_hit_rate.pct": "icc_hit_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
        "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    out, traffic_out, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    res = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"kernel": r[hdr.index("Kernel Name")], "id": r[hdr.index("ID")], "report": rep}
            for m, k in WANT.items():
                if m in hdr:
                    i = hdr.index(m)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    d[k] = v * UNIT.get(units[i], 1.0) if k in ("duration", "dram_read", "dram_write") else v
            res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:60], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k not in ("kernel",)})
    if traffic_out:
        tr = {"source": "ncu --set full --clock-control none, python bench.py --steps 1 (" + out + ")"}
        tc = [d for d in res if "k_als_rows_tc64" in d["kernel"]]
        sv = [d for d in res if "k_svm_sweep_warp" in d["kernel"]]
        if len(tc) >= 2:
            tr["als_half_step_users"] = tc[0]["dram_read"] + tc[0]["dram_write"]
            tr["als_half_step_items"] = tc[1]["dram_read"] + tc[1]["dram_write"]
        if sv:
            tr["svm_block_sweep"] = sv[0]["dram_read"] + sv[0]["dram_write"]
        json.dump(tr, open(traffic_out, "w"), indent=1)
        print(tr)


if __name__ == "__main__":
    main()
