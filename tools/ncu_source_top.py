"""Top source lines by executed instructions from an .ncu-rep captured with --import-source on.

    python tools/ncu_source_top.py gpurun_out/prof.ncu-rep [file-suffix] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    suffix = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            secs.append(cur)
        elif r[0] == "Line No" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur:
            cur["rows"].append(r)
    grand = 0
    for s in secs:
        ie = s["hdr"].index("Instructions Executed")
        s["tot"] = sum(int(r[ie]) for r in s["rows"] if r[0].isdigit() and r[ie].isdigit())
        grand += s["tot"]
    for s in secs:
        print("== %s: %d warp instructions (%.1f %%)" % (s["file"], s["tot"], 100.0 * s["tot"] / max(grand, 1)))
    for s in secs:
        if suffix and not s["file"].endswith(suffix):
            continue
        h = s["hdr"]
        ie, ist = h.index("Instructions Executed"), h.index("# Samples")
        lines = [(int(r[ie]), int(r[ist]) if r[ist].isdigit() else 0, r[0], r[1]) for r in s["rows"]
                 if r[0].isdigit() and r[ie].isdigit()]
        print("-- top lines of", s["file"])
        for n, smp, ln, src in sorted(lines, reverse=True)[:top]:
            print("%12d %5.1f%%  samples %6d  L%-4s %s" % (n, 100.0 * n / max(grand, 1), smp, ln, src.strip()[:100]))


if __name__ == "__main__":
    main()
