"""Per-source-line hot spots of one .ncu-rep (captured with --import-source on, code built with -lineinfo).
usage: python tools/ncu_source.py gpurun_out/x.ncu-rep [top-n] [kernel-substring]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, func, hdr = "", "", None
    agg = {}  # (func, file, line) -> [samples, inst, source, stalls dict]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            func = r[1][:60]; continue
        if r[0] == "Line No":
            hdr = r; ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        if want and want not in func:
            continue
        try:
            s = int(r[ix["# Samples"]]); inst = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        key = (func, fpath, r[0])
        e = agg.setdefault(key, [0, 0, r[1].strip(), {}])
        e[0] += s; e[1] += inst
        for c in stall_cols:
            try:
                e[3][c[6:]] = e[3].get(c[6:], 0) + int(r[ix[c]] or 0)
            except ValueError:
                pass
    funcs = sorted({k[0] for k in agg})
    for f in funcs:
        items = [(v, k) for k, v in agg.items() if k[0] == f]
        tot = sum(v[0] for v, _ in items)
        toti = sum(v[1] for v, _ in items)
        print(f"== {f}  samples={tot} warp-instructions={toti}")
        items.sort(key=lambda x: -x[0][0])
        for v, k in items[:top]:
            st = sorted(((n, c) for c, n in v[3].items() if n), reverse=True)[:3]
            print(f"{100.0 * v[0] / max(tot, 1):5.1f}% inst={100.0 * v[1] / max(toti, 1):4.1f}% {k[1]}:{k[2]:>4} {' '.join(f'{c}:{n}' for n, c in st):42s} | {v[2][:120]}")


if __name__ == "__main__":
    main()
