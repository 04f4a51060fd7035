"""Per-source-line instruction / stall-sample shares from an .ncu-rep source page (kernels compiled with -lineinfo).
usage: python scripts/ncu_lines.py <both.csv from `ncu -i rep --page source --csv --print-source cuda,sass`> <source file> [top]"""
import collections, csv, sys

def main(path, srcfile, top=30, kernel=""):
    rows = list(csv.reader(open(path)))
    fn = [i for i, r in enumerate(rows) if r and r[0] == "Function Name" and kernel in r[1]]
    first = fn[0] if fn else 0
    start = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and i > first][0]
    ends = [i for i, r in enumerate(rows) if r and r[0] == "Function Name" and i > start]
    end = ends[0] if ends else len(rows)
    hdr = rows[start]
    iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in rows[start + 1:end]:
        if len(r) > iT and r[0].isdigit():
            try:
                a = agg[int(r[0])]; a[0] += int(r[iS]); a[1] += int(r[iI]); a[2] += int(r[iT])
            except ValueError:
                pass
    totS, totI = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    src = open(srcfile).read().split("\n")
    print(f"main section: {totI/1e6:.1f} M warp instructions, {totS} samples")
    marks = [(i + 1, l.strip()) for i, l in enumerate(src) if "// ----" in l or "__global__" in l or "__device__" in l]
    marks.append((len(src) + 1, "EOF"))
    for (a, name), (b, _) in zip(marks, marks[1:]):
        sel = [v for k, v in agg.items() if a <= k < b]
        if sel and sum(x[1] for x in sel) > 0.002 * totI:
            I, S, T = sum(x[1] for x in sel), sum(x[0] for x in sel), sum(x[2] for x in sel)
            print(f"  {100*I/totI:5.1f}% inst {100*S/totS:5.1f}% smp  thr {T/max(I,1):4.1f}  L{a}-{b-1} {name[:90]}")
    print("top lines by samples:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  L{k:4d} {100*v[0]/totS:5.1f}% smp {100*v[1]/totI:5.1f}% inst thr {v[2]/max(v[1],1):4.1f}  {src[k-1].strip()[:100]}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30, sys.argv[4] if len(sys.argv) > 4 else "")
