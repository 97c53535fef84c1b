"""Top source lines of a kernel by warp-stall samples, from an `ncu --set full --import-source on` report (-lineinfo build):
    python scripts/ncu_source_hotspots.py report.ncu-rep [top_n] > profiles/<name>_source_hotspots.csv
Reads `ncu -i report --page source --csv --print-source cuda,sass` (no GPU needed) and keeps the per-source-line rows."""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    path, hdr, rows, kernel = None, None, [], None
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1]; continue
        if r[0] == "Function Name":
            kernel = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and r[0].isdigit() and len(r) >= len(hdr) - 1:
            d = dict(zip(hdr[4:], r[4:]))          # columns 0..3 are Line No, Source, Address, Source (SASS)
            num = lambda k: int(d.get(k, "0")) if d.get(k, "0").isdigit() else 0
            rows.append((path.split("/csrc/")[-1], int(r[0]), r[1].strip(), num("Warp Stall Sampling (All Samples)"),
                         num("Warp Stall Sampling (Not-issued Samples)"), num("Instructions Executed"),
                         num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive")))
    tot = sum(x[3] for x in rows) or 1
    tot_i = sum(x[5] for x in rows) or 1
    w = csv.writer(sys.stdout)
    w.writerow(["# " + (kernel or ""), f"stall samples {tot}", f"warp instructions {tot_i}"])
    w.writerow(["file:line", "samples_share", "not_issued_share", "instr_share", "smem_wavefronts", "smem_excess_wavefronts", "source"])
    for f, ln, src, s, ni, ins, wf, wfx in sorted(rows, key=lambda x: -x[3])[:top]:
        w.writerow([f"{f}:{ln}", f"{s / tot:.4f}", f"{ni / tot:.4f}", f"{ins / tot_i:.4f}", wf, wfx, src[:140]])


if __name__ == "__main__":
    main()
