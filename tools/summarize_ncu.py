"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text/JSON file for profiles/.
    python tools/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
        for d in res:
            f.write("\n" + d["kernel"][:160] + "\n")
            for k, v in d.items():
                if k != "kernel":
                    f.write(f"  {k:75s} {v}\n")
    print(out, len(res), "kernels")
    return res


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
