"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [out.txt] [header line ...]"""
import collections, csv, sys


def main(path, out=None, header=()):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        a = agg[r[kn][:120]]
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = list(header) + [f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{a[1]:10.1f} us {a[0]:4d} {100 * a[1] / tot:5.1f}%  {k}")
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3:])
