"""Per-role stall breakdown of a `ncu --set full --import-source on` capture of vq_tc_kernel.

    python scripts/ncu_roles.py gpurun_out/prof_XXX.ncu-rep

Exports the source page (`ncu -i ... --page source --csv`), splits the SASS into the warp roles of the kernel by
landmark instructions (UBLKCP: producer / streamer, UTCHMMA: MMA issuer, F2FP: converter, LDTM: filter, STG.E.EF.128:
gather) and prints, per role, the warp-state samples by stall reason plus the hottest instructions.  This is the
analysis behind the "what bounds it" paragraphs of profiles/README.md."""
import csv
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    src = [r[col["Source"]] for r in data]
    samp = [int(r[col["# Samples"]] or 0) for r in data]
    inst = [int(r[col["Instructions Executed"]] or 0) for r in data]
    tot = max(1, sum(samp))

    def where(pat):
        return [i for i, x in enumerate(src) if pat in x]

    def agg(lo, hi_, name):
        lo, hi_ = max(0, lo), min(len(data), hi_)
        s = sum(samp[lo:hi_])
        d = {h: sum(int(r[col[h]] or 0) for r in data[lo:hi_]) for h in stall}
        top = sorted(d.items(), key=lambda kv: -kv[1])[:7]
        print("%-12s [%5d,%5d) samples %7d (%5.1f%%) inst %8.1fM  %s" % (
            name, lo, hi_, s, 100.0 * s / tot, sum(inst[lo:hi_]) / 1e6, " ".join("%s=%d" % (k[6:], v) for k, v in top)))

    print("%d instructions, %d samples, %.1fM warp-instructions" % (len(data), tot, sum(inst) / 1e6))
    agg(0, len(data), "all")
    f2fp, ldtm, stg, utc, ublk = where("F2FP"), where("LDTM"), where("STG.E.EF.128"), where("UTCHMMA"), where("UBLKCP")
    if ublk:
        agg(ublk[0] - 40, ublk[-1] + 40, "producer+")
    if utc:
        agg(utc[0] - 120, utc[-1] + 60, "mma")
    if f2fp:
        agg(f2fp[0] - 200, f2fp[-1] + 60, "converter")
    for l in ldtm:
        agg(l - 40, l + 120, "filter@%d" % l)
    if stg:
        agg(stg[0] - 120, stg[-1] + 40, "gather")
    print("hottest instructions:")
    for i in sorted(sorted(range(len(data)), key=lambda i: -samp[i])[:30]):
        d = {h: int(data[i][col[h]] or 0) for h in stall}
        t = sorted(d.items(), key=lambda kv: -kv[1])[:3]
        print("%5d %7d  %-64s %s" % (i, samp[i], src[i][:64], " ".join("%s=%d" % (k[6:], v) for k, v in t)))


if __name__ == "__main__":
    main(sys.argv[1])
