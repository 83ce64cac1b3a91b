"""Key per-kernel metrics of an ncu --set full report: python profiles/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("lts__t_sectors_srcunit_tex.sum", "L2sect_from_L1"),
        ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.max", "cycles")]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void <unnamed>::", "").replace("(sdt_conv_desc)", "")
        parts = ["%-22s grid %-14s" % (name[:22], r[ix["Grid Size"]])]
        for m, short in WANT:
            if m in ix:
                v = r[ix[m]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (short, v, units[ix[m]] if short in ("time", "dram_rd", "dram_wr") else ""))
        print(" ".join(parts))
if __name__ == "__main__":
    main(sys.argv[1])
