"""Per-phase stall breakdown of a barrier-structured kernel from an ncu report's source page: samples are summed between
consecutive BAR.SYNC instructions (a stall_barrier sample is charged to the instruction AFTER the barrier it waited at).
usage: python profiles/phases.py <report.ncu-rep> <out.json> [note]"""
import csv
import json
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else ""
    idx = {h: i for i, h in enumerate(rows[hi])}
    seg, cur = [], None
    total = 0

    def new(addr):
        return {"from": addr, "samples": 0, "stall_barrier": 0, "stall_long_sb": 0, "stall_short_sb": 0, "stall_wait": 0,
                "warp_instructions": 0, "sass_instructions": 0}

    for r in rows[hi + 1:]:
        if len(r) <= idx["# Samples"]:
            continue
        if cur is None:
            cur = new(r[idx["Address"]][-5:])
        n = int(r[idx["# Samples"]] or 0)
        total += n
        cur["samples"] += n
        for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait"):
            cur[k] += int(r[idx[k]] or 0)
        cur["warp_instructions"] += int(r[idx["Instructions Executed"]] or 0)
        cur["sass_instructions"] += 1
        if "BAR.SYNC" in r[idx["Source"]] or "EXIT" in r[idx["Source"]]:
            cur["to"] = r[idx["Address"]][-5:] + " " + r[idx["Source"]].strip().split(";")[0][:28]
            seg.append(cur)
            cur = None
    seg = [s for s in seg if s["samples"] * 200 > total]
    for s in seg:
        s["share_of_samples"] = round(s["samples"] / total, 4)
    json.dump({"kernel": kernel, "total_samples": total, "note": sys.argv[3] if len(sys.argv) > 3 else "", "segments": seg},
              open(out, "w"), indent=1)
    print(json.dumps(seg, indent=1)[:1500])


if __name__ == "__main__":
    main()
