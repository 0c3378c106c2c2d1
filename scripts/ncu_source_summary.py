"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines."""
import csv
import sys
import collections

csv.field_size_limit(10 ** 9)


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    # several kernels may be concatenated: split on 'Kernel Name' rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks[:1] if len(sys.argv) < 4 else blocks:
        hdr = b["rows"][0]
        idx = {h: i for i, h in enumerate(hdr)}
        data = b["rows"][1:]
        tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
        print("kernel:", b["name"][:100], "instructions:", len(data), "samples:", tot)
        stalls = [h for h in hdr if h.startswith("stall_")]
        agg = collections.Counter()
        for r in data:
            for s in stalls:
                agg[s] += int(r[idx[s]] or 0)
        st = sum(agg.values())
        print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(st, 1):.1f}%" for k, v in agg.most_common(10)))
        opagg = collections.Counter()
        execagg = collections.Counter()
        for r in data:
            op = r[idx["Source"]].split()
            op = [o for o in op if not o.startswith("@")]
            name = op[0].split(".")[0] if op else "?"
            opagg[name] += int(r[idx["# Samples"]] or 0)
            execagg[name] += int(r[idx["Instructions Executed"]] or 0)
        print("samples by opcode:", ", ".join(f"{k} {100 * v / max(tot, 1):.1f}%" for k, v in opagg.most_common(14)))
        te = sum(execagg.values())
        print("executed by opcode:", ", ".join(f"{k} {100 * v / max(te, 1):.1f}%" for k, v in execagg.most_common(14)))
        order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:top]
        for i in sorted(order):
            r = data[i]
            top_stall = max(stalls, key=lambda s: int(r[idx[s]] or 0))
            print(f"  {i:5d} {int(r[idx['# Samples']]):6d} {100 * int(r[idx['# Samples']]) / max(tot, 1):5.1f}%  "
                  f"{top_stall[6:]:14s} {r[idx['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
