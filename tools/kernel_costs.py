"""Join the manifest of tools/profile_kernels.py with the ncu CSV of the same run -> profiles/kernel_costs.json.
    python -m tools.kernel_costs <manifest.jsonl> <ncu_raw.csv> <capture name> [out.json]
Per key: executed lane-instructions per unit (smsp__inst_executed x 32 / units), DRAM bytes per unit, kernel time, and the
pipe utilisation ncu saw -- everything bench.py needs for `roofline*.frac`, tied to the sources by bench.source_sha()."""
import csv
import json
import re
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402

PIPES = {"issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "fmaheavy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "l1tex_lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def main():
    manifest = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
    rows = list(csv.reader(open(sys.argv[2])))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, units_row, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    unit_of = dict(zip(hdr, units_row))
    launches = [{"name": r[col["Kernel Name"]], "row": r} for r in data if len(r) == len(hdr)]
    pos = 0
    out = {}
    for m in manifest:
        got = []
        for rx in m["kernels"]:
            # the launches of one case are consecutive; take the next launch whose name matches
            for j in range(pos, len(launches)):
                if re.search(re.escape(rx), launches[j]["name"]):
                    got.append(launches[j])
                    pos = max(pos, j + 1) if rx == m["kernels"][-1] else pos
                    break
        if len(got) != len(m["kernels"]):
            print(f"warning: {m['key']}: matched {len(got)} of {len(m['kernels'])} kernels", file=sys.stderr)
            continue

        def val(l, name):
            v = num(l["row"][col[name]]) if name in col else None
            if v is None:
                return None
            u = unit_of.get(name, "")
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3, "usecond": 1e-3,
                     "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u)
            return v * scale if scale is not None and ("bytes" in name or "time" in name) else v

        inst = sum(val(l, "smsp__inst_executed.sum") or 0 for l in got)
        dram = sum((val(l, "dram__bytes_read.sum") or 0) + (val(l, "dram__bytes_write.sum") or 0) for l in got)
        ms = sum(val(l, "gpu__time_duration.sum") or 0 for l in got)
        main_l = max(got, key=lambda l: val(l, "gpu__time_duration.sum") or 0)
        out[m["key"]] = {"unit": m["unit"], "units_in_capture": m["units"], "warp_inst_x32_per_unit": inst * 32 / m["units"],
                         "dram_bytes_per_unit": dram / m["units"], "ms_in_capture(ncu, cold, serialised)": ms, "kernels": [l["name"][:80] for l in got],
                         "shape": m.get("shape"), "capture": sys.argv[3],
                         **{k: val(main_l, v) for k, v in PIPES.items()}}
    res = {"source_sha": bench.source_sha(), "how": "tools/profile_round.sh: ncu --clock-control none over tools/profile_kernels.py (one launch per case); "
           "numbers taken under ncu are never bench values, only per-unit instruction / byte counts and pipe shares", "kernels": out}
    path = sys.argv[4] if len(sys.argv) > 4 else "profiles/kernel_costs.json"
    json.dump(res, open(path, "w"), indent=1)
    print(f"{path}: {len(out)} kernels, source {res['source_sha']}")


if __name__ == "__main__":
    main()
