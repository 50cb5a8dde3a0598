"""Warp-stall samples of a warp-specialised tcgen05 kernel, split by warp ROLE - development aid.

    python scripts/ncu_stall_by_role.py gpurun_out/r01f_halo.ncu-rep [launch indices ...] [--top N]

Input: an ncu report captured with `--section SourceCounters --import-source on` (see scripts/one_forward.py).  For every
requested launch the SASS listing is exported (`ncu -i ... --page source --csv --print-source sass`) and cut into three
regions by the position of the TMA loads (UTMALDG) and the MMAs (UTCHMMA): prologue + producer | epilogue | MMA issuer -
the order the role branches are laid out in by the compiler for the kernels of this repository.  Printed per region:
samples, the top stall reasons, the hottest instructions; for the epilogue also how many samples sit in the spin on the
accumulator-full barrier (idle) against inside a tile (busy).  Sampling is per warp, so a role's share of ALL samples only
says how many warps it has; where its samples sit is the information.
"""
import csv
import subprocess
import sys
import tempfile


def export(rep, launch):
    with tempfile.NamedTemporaryFile("w+", suffix=".csv") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(launch),
                        "--launch-count", "1"], stdout=f, stderr=subprocess.DEVNULL, check=False)
        f.seek(0)
        rows = list(csv.reader(f))
    if len(rows) < 3:
        return None, None, None
    name = rows[0][1] if len(rows[0]) > 1 else "?"
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].strip().isdigit()]
    # ncu prints the listing once per view it knows; keep the first copy
    first_addr = data[0][0]
    for i in range(1, len(data)):
        if data[i][0] == first_addr:
            data = data[:i]
            break
    return name, hdr, data


def main():
    argv = sys.argv[1:]
    top_n = 6
    if "--top" in argv:
        i = argv.index("--top")
        top_n = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    rep = args[0]
    launches = [int(v) for v in args[1:] if v.isdigit()] or [0]
    for k in launches:
        name, hdr, d = export(rep, k)
        if d is None:
            print(f"launch {k}: no source page in the report")
            continue
        S, E, SRC = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
        r0 = hdr.index("stall_barrier")
        reason_names = hdr[r0:r0 + 17]
        val = lambda r, i: int(r[i]) if r[i].strip().isdigit() else 0
        total = sum(val(r, S) for r in d)
        mma = [i for i, r in enumerate(d) if "UTCHMMA" in r[SRC]]
        tma = [i for i, r in enumerate(d) if "UTMALDG" in r[SRC]]
        print(f"\n== launch {k}: {name[:110]}\n   {total} samples, {len(d)} SASS instructions")
        if not mma or not tma:
            regions = [("whole kernel", 0, len(d))]
        else:
            pb, ib = tma[-1] + 30, max(mma[0] - 60, tma[-1] + 31)
            regions = [("prologue + TMA producer", 0, pb), ("epilogue", pb, ib), ("MMA issuer (+ exit)", ib, len(d))]
        for label, a, b in regions:
            reg = d[a:b]
            s = sum(val(r, S) for r in reg)
            agg = sorted(((sum(val(r, r0 + j) for r in reg), reason_names[j][6:]) for j in range(17)), reverse=True)[:4]
            print(f"  {label:26s} {s:6d} samples ({100.0 * s / max(total, 1):4.1f} %)  " + ", ".join(f"{n} {v}" for v, n in agg))
            if label == "epilogue":
                spin = 0
                for i, r in enumerate(reg):
                    if "TRYWAIT" in r[SRC]:
                        spin += val(r, S) + (val(reg[i + 1], S) if i + 1 < len(reg) else 0)
                print(f"      idle (spinning on a barrier) {spin}, busy inside a tile {s - spin}")
            for r in sorted(reg, key=lambda r: -val(r, S))[:top_n]:
                if val(r, S) == 0:
                    break
                why = max(((val(r, r0 + j), reason_names[j][6:]) for j in range(17)))[1]
                print(f"      {val(r, S):6d}  x{r[E].strip():>8s}  {why:18s} {r[SRC].strip()[:72]}")


if __name__ == "__main__":
    main()
