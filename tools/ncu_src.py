"""Source-level digest of an .ncu-rep (captured with --import-source on): executed warp
instructions and stall samples per contiguous SASS region of equal execution count, plus
the shared-memory loads with excess wavefronts.  Read here, no GPU needed.

    python tools/ncu_src.py gpurun_out/x.ncu-rep regex:kernel_name [min_pct]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    data = [r for r in rows if len(r) > ia and r[ia].isdigit()]
    # the page lists the kernel once per launch captured: keep the first copy
    first = data[0][1]
    for i in range(1, len(data)):
        if data[i][1] == first and data[i][0] != data[0][0]:
            pass
    addrs = {}
    uniq = []
    for r in data:
        if r[0] in addrs:
            break
        addrs[r[0]] = 1
        uniq.append(r)
    data = uniq
    tot = sum(int(r[ia]) for r in data)
    tots = sum(int(r[isamp]) for r in data)
    print("total warp instructions %d, SASS instructions %d, samples %d" % (tot, len(data), tots))
    cur, start, s, ns = None, 0, 0, 0
    for i, r in enumerate(data + [None]):
        c = int(r[ia]) if r else -1
        if cur is None or c != cur:
            if cur is not None and (s > tot * min_pct / 100 or ns > tots * min_pct / 100):
                print("%5d-%5d x%-9d n=%4d  instr %10d (%4.1f%%)  samples %6d (%4.1f%%)  %s" % (
                    start, i - 1, cur, i - start, s, 100.0 * s / tot, ns, 100.0 * ns / max(tots, 1),
                    data[start][isrc].strip()[:40]))
            cur, start, s, ns = c, i, 0, 0
        if r:
            s += c
            ns += int(r[isamp])
    hot = sorted(((int(r[isamp]), i, r) for i, r in enumerate(data)), reverse=True)[:12]
    print("hottest instructions by stall samples:")
    for n, i, r in hot:
        print("  %5d %6d  %s" % (i, n, r[isrc].strip()[:70]))
    for i, r in enumerate(data):
        w, wi = int(r[iw] or 0), int(r[iwi] or 0)
        if wi and w > 1.5 * wi and w > tot / 400:
            print("  smem %5d %-50s wavefronts %d ideal %d" % (i, r[isrc].strip()[:50], w, wi))


if __name__ == "__main__":
    main()
