"""Summarise one kernel of an `ncu --set full --import-source on` report the way profiles/r1_*_summary.txt do:
pipe utilisation, stall cycles per issued instruction, warp-instructions per scan element, and — from the SASS source
page — where the warp-stall samples fall, split by how often an instruction executes (state loop / per-chunk code /
per-kernel code) and by opcode inside the hot loop.

    python scripts/ncu_report.py gpurun_out/scan_v3.ncu-rep [--elements 4294967296] [--top 25] [--dump-loop]

Needs `ncu` on PATH (reads the report with `ncu -i ... --page raw|source --csv`); no GPU required."""
import argparse
import collections
import csv
import io
import re
import subprocess


def _page(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def _f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--elements", type=float, default=4 * 512 * 131072 * 16,
                    help="(token, channel, state) elements of the launch (default: Caduceus-PS headline shapes)")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--dump-loop", action="store_true", help="print the hottest loop instruction by instruction")
    a = ap.parse_args()

    raw = _page(a.report, "raw")
    d = {h: v for h, v in zip(raw[0], raw[2])}
    print("kernel:", d.get("Kernel Name", "?"))
    units = {h: u for h, u in zip(raw[0], raw[1])}
    print(f"duration {_f(d['gpu__time_duration.sum']):.4f} {units.get('gpu__time_duration.sum', 'ms')}   grid {d['launch__grid_size']} x {d['launch__block_size']} threads, "
          f"{d['launch__registers_per_thread']} regs")
    print(f"issue slots {_f(d['smsp__issue_active.avg.pct_of_peak_sustained_active']):.1f} %   "
          f"MUFU {_f(d['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']):.1f} %   "
          f"FMA {_f(d['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']):.1f} %   "
          f"ALU {_f(d['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']):.1f} %   "
          f"LSU {_f(d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']):.1f} %   "
          f"smem wavefronts {_f(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']):.1f} %")
    print(f"warps active {_f(d['smsp__warps_active.avg.per_cycle_active']):.2f} / eligible "
          f"{_f(d['smsp__warps_eligible.avg.per_cycle_active']):.2f} per scheduler")
    print(f"DRAM read {d['dram__bytes_read.sum']} + write {d['dram__bytes_write.sum']}")
    inst = _f(d["smsp__inst_executed.sum"])
    print(f"warp-instructions per element {inst / (a.elements / 32):.2f}")
    stalls = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), _f(v)) for h, v in d.items()
              if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    print("stall cycles per issued instruction: " + "  ".join(f"{n} {v:.2f}" for n, v in sorted(stalls, key=lambda t: -t[1])[:9]))

    src = _page(a.report, "source", ("--print-source", "sass"))
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    S, W, E = ix["Source"], ix["Warp Stall Sampling (All Samples)"], ix["Instructions Executed"]
    rows = [(r[S].strip(), int(r[W] or 0), int(r[E] or 0)) for r in src[2:] if len(r) > E]
    tot = sum(w for _, w, _ in rows) or 1
    mx = max(e for _, _, e in rows) or 1
    classes = [("state loop (>= 40 % of the max execution count)", lambda e: e >= 0.4 * mx),
               ("per-chunk code", lambda e: 0.02 * mx <= e < 0.4 * mx), ("rest", lambda e: e < 0.02 * mx)]
    itot = sum(e for _, _, e in rows) or 1
    for name, pred in classes:
        w = sum(w for _, w, e in rows if pred(e))
        i = sum(e for _, _, e in rows if pred(e))
        print(f"  {100 * w / tot:5.1f} % of the stall samples, {100 * i / itot:5.1f} % of the executed instructions: {name}")
    by_op = collections.Counter()
    n_op = collections.Counter()
    for s, w, e in rows:
        if e >= 0.4 * mx:
            op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
            by_op[op] += w
            n_op[op] += 1
    print("state loop, samples by opcode: " + "  ".join(f"{op} {100 * w / tot:.1f}% ({n_op[op]})" for op, w in by_op.most_common(10)))
    print(f"top {a.top} instructions by stall samples:")
    for s, w, e in sorted(rows, key=lambda t: -t[1])[:a.top]:
        print(f"  {w:7d}  x{e:<9d} {s[:90]}")
    if a.dump_loop:
        idx = [i for i, (_, _, e) in enumerate(rows) if e >= 0.4 * mx]
        runs, cur = [], [idx[0]]
        for i in idx[1:]:
            if i == cur[-1] + 1:
                cur.append(i)
            else:
                runs.append(cur)
                cur = [i]
        runs.append(cur)
        best = max(runs, key=lambda r: sum(rows[i][1] for i in r))
        print(f"hottest loop: {len(best)} instructions, {sum(rows[i][1] for i in best)} samples")
        for i in best:
            print(f"  {rows[i][1]:6d}  {rows[i][0][:100]}")


if __name__ == "__main__":
    main()
