"""Summarise `ncu --set full` reports (read here, no GPU needed) into the CSV kept under profiles/:
    python tools/ncu_summary.py gpurun_out/<tag>_teacher_gemm.ncu-rep gpurun_out/<tag>_student_kernels.ncu-rep ... > profiles/<tag>_ncu_full_kernels.csv
Columns: duration, tensor pipe % of peak (elapsed), issue slots %, DRAM read / write MB, L2 -> SM MB, shared-memory LSU wavefronts %,
registers per thread, SM clock during the capture, grid size."""
import csv
import io
import os
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", 1.0),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%", 1.0), ("dram__bytes_read.sum", "rdMB", None),
        ("dram__bytes_write.sum", "wrMB", None), ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2smMB", None),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu%", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("sm__cycles_elapsed.avg.per_second", "GHz", None), ("launch__grid_size", "grid", 1.0)]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def main():
    out = csv.writer(sys.stdout)
    out.writerow(["capture", "kernel"] + [c[1] for c in COLS])
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        cap = os.path.basename(rep).replace(".ncu-rep", "").split("_", 1)[-1]
        for r in rows[2:]:
            vals = []
            for name, short, _ in COLS:
                i = idx[name]
                v = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0) if r[i] not in ("", "n/a") else float("nan")
                if short == "us":
                    v *= 1e6
                elif short in ("rdMB", "wrMB", "l2smMB"):
                    v /= 1e6
                elif short == "GHz":
                    v /= 1e9
                vals.append(f"{v:.6f}" if short not in ("regs", "grid") else str(int(v)))
            out.writerow([cap, r[idx["Kernel Name"]][:100]] + vals)


if __name__ == "__main__":
    main()
