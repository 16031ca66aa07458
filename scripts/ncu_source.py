"""Per-CUDA-source-line stall/instruction summary of an .ncu-rep (needs -lineinfo and --import-source on):
python scripts/ncu_source.py REP kernel_regex [top] [launch_index]"""
import csv
import io
import os
import subprocess
import sys


def fl(v):
    try:
        return float(v)
    except (TypeError, ValueError):
        return 0.0


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    recs, cur_file, hdr, launch, seen_fn = [], "", None, -1, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            if cur_file and (seen_fn is None or cur_file in seen_fn):
                if seen_fn is None:
                    seen_fn = set()
            continue
        if hdr is None or not r[0].isdigit():
            continue
        recs.append((cur_file, int(r[0]), r[1], dict(zip(hdr[4:], r[4:]))))
    # several launches repeat the same (file,line) keys; keep launch `which` by counting repeats
    count, picked = {}, []
    for f, ln, src, d in recs:
        k = (f, ln)
        c = count.get(k, 0)
        count[k] = c + 1
        if c == which:
            picked.append((f, ln, src, d))
    tot_s = sum(fl(d.get("# Samples")) for *_, d in picked)
    tot_i = sum(fl(d.get("Instructions Executed")) for *_, d in picked)
    print(f"total samples {tot_s:.0f}   total warp-instructions {tot_i:.0f}")
    stall_cols = [c for c in (hdr or []) if c.startswith("stall_") and "Not Issued" not in c]
    for f, ln, src, d in sorted(picked, key=lambda t: -fl(t[3].get("# Samples")))[:top]:
        s = fl(d.get("# Samples"))
        i = fl(d.get("Instructions Executed"))
        st = sorted(((fl(d.get(c)), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{s:7.0f} {100*s/max(tot_s,1):5.1f}% inst {100*i/max(tot_i,1):5.1f}%  "
              f"{' '.join(f'{n}:{v:.0f}' for v, n in st if v > 0):38s} {f}:{ln}: {src.strip()[:90]}")


if __name__ == "__main__":
    main()
